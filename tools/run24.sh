#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "four_step" > gpurun_out/pytest24.txt 2>&1
tail -15 gpurun_out/pytest24.txt
python tools/sweep.py n16384 n16384:::FSEA_FOUR_OLD=1 n32768 n32768:::FSEA_FOUR_OLD=1 n65536 n65536:::FSEA_FOUR_OLD=1 cfg4 cfg4:::FSEA_FOUR_OLD=1 cfg5 cfg5:::FSEA_FOUR_OLD=1 > gpurun_out/sweep24.txt 2>&1
cat gpurun_out/sweep24.txt
ncu --set full --clock-control none --import-source on -k regex:fsea_ -s 6 -c 2 -o gpurun_out/prof24_65536 python tools/sweep.py n65536 > gpurun_out/ncu24.log 2>&1
ls -la gpurun_out/prof24*
