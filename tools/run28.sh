#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "four_step" > gpurun_out/pytest28.txt 2>&1
tail -15 gpurun_out/pytest28.txt
python tools/sweep.py n16384 n16384:::FSEA_FOUR_SPLIT=1 n32768 n32768:::FSEA_FOUR_SPLIT=1 n65536 n65536:::FSEA_FOUR_SPLIT=1 cfg4 cfg5 > gpurun_out/sweep28.txt 2>&1
cat gpurun_out/sweep28.txt
ncu --set full --clock-control none --import-source on -k regex:fsea_ -s 3 -c 1 -o gpurun_out/prof28_16384 python tools/sweep.py n16384 > gpurun_out/ncu28.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fsea_ -s 3 -c 1 -o gpurun_out/prof28_65536 python tools/sweep.py n65536 > gpurun_out/ncu28b.log 2>&1
ls -la gpurun_out/prof28*
