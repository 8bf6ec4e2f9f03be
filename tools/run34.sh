#!/bin/bash
# frame-per-unit ("half") kernels for 2048/4096/8192: tests, A/B sweep, ncu summaries
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest34.txt 2>&1
tail -5 gpurun_out/pytest34.txt
python tools/sweep.py cfg2 n2048:f32x2h n2048:f32x2 n4096:f32x2h n4096:f32x2 n8192:f32x2h n8192:f32x2 cfg3:f32x2h cfg3:f32x2 \
   n16384 n32768 n65536 n65536:::FSEA_FOUR_ALT=1 > gpurun_out/sweep34.txt 2>&1
cat gpurun_out/sweep34.txt
SWEEP_STEPS=1 SWEEP_WARMUP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fsea_ -c 6 -o gpurun_out/prof34 \
   python tools/sweep.py n2048:f32x2h n8192:f32x2h n4096:f32 > gpurun_out/ncu34.log 2>&1
python tools/ncu_summary.py gpurun_out/prof34.ncu-rep > gpurun_out/ncu34_summary.txt 2>&1
sz=$(stat -c %s gpurun_out/prof34.ncu-rep); if [ $sz -gt 45000000 ]; then rm gpurun_out/prof34.ncu-rep; fi
cat gpurun_out/ncu34_summary.txt
