#!/bin/bash
# Round-end evidence run (under gpurun, 1 GPU): tests, bench (headline + secondary), reference arm, launch list, full ncu
# captures of every kernel family, FIR / operator benches, sanitizer.  Everything lands in gpurun_out/final/ ; what is
# to be judged is copied into profiles/r02_final/.
O=gpurun_out/final
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1
tail -3 $O/pytest_gpu.txt
python bench.py > $O/bench.json 2> $O/bench.err
python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-secondary > $O/launch_bench.log 2>&1
# one full capture of the headline kernel inside bench.py
ncu --set full --clock-control none --import-source on -k regex:fsea_fft -s 3 -c 1 -o $O/prof_n1024 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-secondary > $O/ncu_bench.log 2>&1
python tools/ncu_summary.py $O/prof_n1024.ncu-rep > $O/ncu_full_n1024_f32x2.txt 2>&1
ncu -i $O/prof_n1024.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,json
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[2]
g=lambda k: float(v[h.index(k)].replace(',',''))
u=lambda k: rows[1][h.index(k)]
def b(k):
    x=g(k); un=u(k)
    return x*{'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}[un]
print(json.dumps({'kernel': v[h.index('Kernel Name')], 'fft_size': 1024, 'frames': 262144, 'dram_bytes_read': b('dram__bytes_read.sum'), 'dram_bytes_write': b('dram__bytes_write.sum'), 'dram_bytes_per_launch': b('dram__bytes_read.sum')+b('dram__bytes_write.sum'), 'source': 'ncu --set full, one launch of the bench batch (round 2 final run)'}))
" > $O/ncu_traffic.json 2>/dev/null
rm -f $O/prof_n1024.ncu-rep
# one full capture per other kernel family
for w in n2048@65536 n4096@32768 n8192@16384 cfg3@16384 n16384@8192 cfg4@8192 n32768@4096 n65536@2048; do
  tag=$(echo $w | tr '@' '_')
  SWEEP_STEPS=1 SWEEP_WARMUP=1 timeout 300 ncu --set full --clock-control none -k regex:fsea_ -s 1 -c 1 -o $O/prof_$tag python tools/sweep.py $w > $O/ncu_$tag.log 2>&1
  python tools/ncu_summary.py $O/prof_$tag.ncu-rep > $O/ncu_full_$tag.txt 2>&1
  rm -f $O/prof_$tag.ncu-rep
done
python tools/sweep.py cfg2@65536:f64 n8192@8192:f64 cfg4@2048:f64 n65536@256:f64 2>&1 | cut -c1-200 > $O/f64_mode.txt
python tools/bench_fir.py --taps 51 > $O/bench_fir_51.json 2> $O/bench_fir_51.err
python tools/bench_fir.py --taps 97 > $O/bench_fir_97.json 2> $O/bench_fir_97.err
python tools/bench_operator.py > $O/bench_operator_per_draw.json 2> $O/bench_operator.err
python tools/sweep_mgpu.py --gpus 1 --freqs 1000 --rows 512 --mode p2p > $O/sweep_job_1gpu.json 2> $O/sweep_job_1gpu.err
for tool in memcheck racecheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool python tools/sanitize.py > $O/compute_sanitizer_$tool.txt 2>&1; echo "$tool rc=$?" >> $O/compute_sanitizer_$tool.txt
done
python - <<'PY'
import json
d = json.loads(open("gpurun_out/final/bench.json").read().strip().splitlines()[-1])
print("headline", round(d["value"]/1e6,1), "Mframes/s frac", round(d["roofline"]["frac"],3), "burst", round(d["roofline"]["burst_frac"],3), "e2e", round(d["e2e"]["value"]/1e6,2), "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None, "s", d["bench_seconds"])
for k, v in d["secondary"].items():
    if v and "roofline" in v: print(k, round(v["value"]/1e6,2), "frac", round(v["roofline"]["frac"],3), v["kernel"][:30])
    else: print(k, json.dumps(v)[:300])
PY
