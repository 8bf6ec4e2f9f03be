O=gpurun_out/r02r; mkdir -p $O
timeout 400 python tools/sweep.py n8192:::FSEA_WS=1 n8192:::FSEA_WS=0 cfg3:::FSEA_WS=1 cfg3:::FSEA_WS=0 n8192:::FSEA_WS=1 n8192:::FSEA_WS=0 cfg3:::FSEA_WS=1 cfg3:::FSEA_WS=0 n16384 > $O/sweep.txt 2>&1; cat $O/sweep.txt | cut -c1-30,42-90,150-250
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "four_step or large_batch or hann or all_sizes" > $O/pytest_gpu.txt 2>&1; tail -2 $O/pytest_gpu.txt
