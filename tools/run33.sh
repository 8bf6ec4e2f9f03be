#!/bin/bash
# ncu --set full of the fused four-step kernels and the 8192 direct kernel; summaries made on the box
mkdir -p gpurun_out
SWEEP_STEPS=1 SWEEP_WARMUP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fsea_ -c 8 -o gpurun_out/prof33 \
   python tools/sweep.py n16384 n65536 n8192 n8192:::FSEA_8192_DIRECT=1 > gpurun_out/ncu33.log 2>&1
python tools/ncu_summary.py gpurun_out/prof33.ncu-rep > gpurun_out/ncu33_summary.txt 2>&1
python tools/ncu_ops.py gpurun_out/prof33.ncu-rep > gpurun_out/ncu33_ops.txt 2>&1
ncu -i gpurun_out/prof33.ncu-rep --page details > gpurun_out/ncu33_details.txt 2>&1
ls -la gpurun_out/
sz=$(stat -c %s gpurun_out/prof33.ncu-rep); if [ $sz -gt 45000000 ]; then rm gpurun_out/prof33.ncu-rep; fi
cat gpurun_out/ncu33_summary.txt
