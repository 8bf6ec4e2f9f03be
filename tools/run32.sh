#!/bin/bash
# generalised four-step family (N = 1024 x N2) + scalar 64-point-per-thread shapes: tests, A/B sweep, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest32.txt 2>&1
tail -5 gpurun_out/pytest32.txt
python tools/sweep.py cfg2 n2048 n2048:f32 n4096 n4096:f32 \
   n8192 n8192:::FSEA_8192_DIRECT=1 cfg3 cfg3:::FSEA_8192_DIRECT=1 \
   n16384 n16384:::FSEA_FOUR_ALT=1 n32768 n32768:::FSEA_FOUR_ALT=1 n65536 n65536:::FSEA_FOUR_ALT=1 cfg4 cfg5 > gpurun_out/sweep32.txt 2>&1
cat gpurun_out/sweep32.txt
SWEEP_STEPS=1 SWEEP_WARMUP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fsea_ -c 12 -o gpurun_out/prof32 \
   python tools/sweep.py n8192 n16384 n65536 n4096:f32 > gpurun_out/ncu32.log 2>&1
ls -la gpurun_out/prof32*
