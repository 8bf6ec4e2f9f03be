#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "scope" > gpurun_out/pytest39.txt 2>&1; tail -2 gpurun_out/pytest39.txt
for tool in memcheck racecheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool python tools/sanitize.py > gpurun_out/sanitizer_$tool.txt 2>&1; echo "$tool rc=$?" >> gpurun_out/sanitizer_$tool.txt
  tail -3 gpurun_out/sanitizer_$tool.txt
done
