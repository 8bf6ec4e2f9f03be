#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "file or iq_filter or golden" > gpurun_out/pytest42.txt 2>&1; tail -2 gpurun_out/pytest42.txt
for w in cfg5 cfg3 cfg4; do
python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_${w}_file.json 2> gpurun_out/bench_${w}_file.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_${w}_file.json').read().strip().splitlines()[-1]); print(d['value'], d['file_ingest'])"
done
SWEEP_STEPS=1 SWEEP_WARMUP=1 timeout 600 ncu --set full --clock-control none -k regex:fsea_ -c 6 -o gpurun_out/prof42 \
   python tools/sweep.py n32768 n65536 > gpurun_out/ncu42.log 2>&1
python tools/ncu_summary.py gpurun_out/prof42.ncu-rep > gpurun_out/ncu42_summary.txt 2>&1
rm -f gpurun_out/prof42.ncu-rep
grep "kernel:\|duration\|DRAM\|FMA\|L1/shared data\|L2 (lts) thr\|stalls" gpurun_out/ncu42_summary.txt | tail -16
