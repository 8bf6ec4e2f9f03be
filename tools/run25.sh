#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "four_step" > gpurun_out/pytest25.txt 2>&1
tail -5 gpurun_out/pytest25.txt
python tools/sweep.py n16384 n32768 n65536 n65536::build_alt/libfsea_tws.so cfg4 cfg5 > gpurun_out/sweep25.txt 2>&1
cat gpurun_out/sweep25.txt
ncu --set full --clock-control none --import-source on -k regex:fsea_ -s 6 -c 2 -o gpurun_out/prof25_65536 python tools/sweep.py n65536 > gpurun_out/ncu25.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fsea_ -s 6 -c 2 -o gpurun_out/prof25_16384 python tools/sweep.py n16384 > gpurun_out/ncu25b.log 2>&1
ls -la gpurun_out/prof25*
