#!/bin/bash
# frame-per-unit kernels with 3 units per block sharing the pass-2 twiddle table in shared memory (g3) vs per-block L1/L2 twiddles
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest37.txt 2>&1
tail -3 gpurun_out/pytest37.txt
python tools/sweep.py n8192 n8192::build_alt/libfsea_g3.so cfg3 cfg3::build_alt/libfsea_g3.so \
   n4096 n4096::build_alt/libfsea_g3.so n2048 n2048::build_alt/libfsea_g3.so > gpurun_out/sweep37.txt 2>&1
cat gpurun_out/sweep37.txt
