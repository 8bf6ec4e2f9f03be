#!/bin/bash
mkdir -p gpurun_out
python tools/sweep.py n2048 n2048:f32 n2048::build_alt/libfsea_s1.so n2048:f32:build_alt/libfsea_s1.so \
  n4096 n4096:f32 n8192 n8192:f32 n8192::build_alt/libfsea_s1.so n8192:f32:build_alt/libfsea_s1.so \
  cfg3 cfg3:f32 cfg3::build_alt/libfsea_s1.so n16384 n16384:f32 n32768 n65536 n65536:f32 cfg4 cfg5 cfg2 cfg2:f32 > gpurun_out/sweep23.txt 2>&1
cat gpurun_out/sweep23.txt
ncu --set full --clock-control none --import-source on -k regex:fsea_fft -s 3 -c 1 -o gpurun_out/prof23_8192 python tools/sweep.py n8192 > gpurun_out/ncu23a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fsea_fft -s 3 -c 1 -o gpurun_out/prof23_2048 python tools/sweep.py n2048 > gpurun_out/ncu23b.log 2>&1
ls -la gpurun_out/*.ncu-rep
