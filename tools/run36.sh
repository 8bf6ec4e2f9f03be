#!/bin/bash
# four-step: row-pair interleaved intermediate for N2 = 64 + streaming (evict-first) spectrum stores; A/B vs HEAD~ (base) and no-stream (ns)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest36.txt 2>&1
tail -3 gpurun_out/pytest36.txt
python tools/sweep.py n65536 n65536::build_alt/libfsea_ns.so n65536::build_alt/libfsea_base.so \
   n32768 n32768::build_alt/libfsea_base.so n16384 n16384::build_alt/libfsea_base.so \
   cfg5 cfg5::build_alt/libfsea_base.so cfg4 cfg4::build_alt/libfsea_base.so n4096 n8192 cfg3 > gpurun_out/sweep36.txt 2>&1
cat gpurun_out/sweep36.txt
