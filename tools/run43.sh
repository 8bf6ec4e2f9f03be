#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/sweep_demo.py --freqs 40 --rows 64 --n 16384 > gpurun_out/sweep_demo_2gpu.json 2> gpurun_out/sweep_demo_2gpu.err; tail -1 gpurun_out/sweep_demo_2gpu.json; tail -3 gpurun_out/sweep_demo_2gpu.err
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_cfg2_2gpu.json 2> gpurun_out/bench_cfg2_2gpu.err; tail -c 900 gpurun_out/bench_cfg2_2gpu.json
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err; tail -c 300 gpurun_out/bench_ref_2gpu.json
timeout 300 $TR --master-port 29514 bench.py --gpus 2 --workload cfg5 --steps 5 --warmup 3 > gpurun_out/bench_cfg5_2gpu.json 2> gpurun_out/bench_cfg5_2gpu.err; tail -c 700 gpurun_out/bench_cfg5_2gpu.json
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sweep" > gpurun_out/pytest43.txt 2>&1; tail -2 gpurun_out/pytest43.txt
