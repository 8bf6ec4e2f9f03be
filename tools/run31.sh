#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest31.txt 2>&1
tail -5 gpurun_out/pytest31.txt
python tools/sweep.py n16384 n32768 n65536 cfg4 cfg5 cfg2 n2048 > gpurun_out/sweep31.txt 2>&1
cat gpurun_out/sweep31.txt
