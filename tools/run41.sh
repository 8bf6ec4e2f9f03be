#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest41.txt 2>&1; tail -3 gpurun_out/pytest41.txt
python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/bench_cfg5_file.json 2> gpurun_out/bench_cfg5_file.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg5_file.json').read().strip().splitlines()[-1]); print(d['value'], d['file_ingest'])"
python bench.py --workload cfg3 --steps 5 --warmup 3 > gpurun_out/bench_cfg3_file.json 2> gpurun_out/bench_cfg3_file.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_file.json').read().strip().splitlines()[-1]); print(d['value'], d['file_ingest'])"
