#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "iq_filter or scope" > gpurun_out/pytest40.txt 2>&1; tail -2 gpurun_out/pytest40.txt
python tools/bench_fir.py --taps 51 > gpurun_out/bench_fir_51.json 2> gpurun_out/bench_fir_51.err; cat gpurun_out/bench_fir_51.json; tail -2 gpurun_out/bench_fir_51.err
python tools/bench_fir.py --taps 97 > gpurun_out/bench_fir_97.json 2> gpurun_out/bench_fir_97.err; cat gpurun_out/bench_fir_97.json
