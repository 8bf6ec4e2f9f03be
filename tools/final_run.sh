#!/bin/bash
# Round-end evidence run (under gpurun): tests, headline bench, launch list, one full ncu capture, size ladder.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1
tail -3 gpurun_out/pytest_gpu.txt
python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_cfg2_reference.json 2> gpurun_out/bench_cfg2_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fsea_fft -s 3 -c 1 -o gpurun_out/prof_final_1024 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
for w in n128 n256 n512 n2048 n4096 n8192 n16384 n32768 n65536 cfg3 cfg4 cfg5; do
  timeout 150 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(d["value"] / 1e6, 2), "Mframes/s", "frac", round(d["roofline"]["frac"], 3) if d.get("roofline") else None,
              "e2e", round(d["e2e"]["value"] / 1e6, 2) if d.get("e2e") else None)
    except Exception as e:
        print(f, "ERR", e)
PY
