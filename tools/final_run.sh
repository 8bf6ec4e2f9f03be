#!/bin/bash
# Round-end evidence run (under gpurun): tests, headline bench, reference arm, launch list, full ncu captures, size ladder.
# Everything lands in gpurun_out/final/ ; copy what is to be judged into profiles/r01_final/.
O=gpurun_out/final
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1
tail -3 $O/pytest_gpu.txt
python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err
python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_cfg2_reference.json 2> $O/bench_cfg2_reference.err
# launch list of the same command (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_bench_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu > $O/launch_bench.log 2>&1
# one full capture of the headline kernel inside bench.py
ncu --set full --clock-control none --import-source on -k regex:fsea_fft -s 3 -c 1 -o $O/prof_n1024 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/ncu_bench.log 2>&1
python tools/ncu_summary.py $O/prof_n1024.ncu-rep > $O/ncu_full_n1024_f32x2.txt 2>&1
# one full capture per other kernel family: frame-per-unit 2048 / 4096 / 8192(+Hann), cluster four-step 16384 / 32768 / 65536
SWEEP_STEPS=1 SWEEP_WARMUP=1 timeout 900 ncu --set full --clock-control none -k regex:fsea_ -c 14 -o $O/prof_family \
   python tools/sweep.py n2048 n4096 n8192 cfg3 n16384 > $O/ncu_family.log 2>&1
python tools/ncu_summary.py $O/prof_family.ncu-rep > $O/ncu_full_family.txt 2>&1
rm -f $O/prof_family.ncu-rep
SWEEP_STEPS=1 SWEEP_WARMUP=1 timeout 600 ncu --set full --clock-control none -k regex:fsea_ -c 6 -o $O/prof_four \
   python tools/sweep.py n32768 n65536 > $O/ncu_four.log 2>&1
python tools/ncu_summary.py $O/prof_four.ncu-rep > $O/ncu_full_four_step_32768_65536.txt 2>&1
rm -f $O/prof_four.ncu-rep
sz=$(stat -c %s $O/prof_n1024.ncu-rep 2>/dev/null || echo 0); if [ $sz -gt 30000000 ]; then rm $O/prof_n1024.ncu-rep; fi
for w in n128 n256 n512 n2048 n4096 n8192 n16384 n32768 n65536 cfg3 cfg4 cfg5; do
  timeout 150 python bench.py --workload $w --steps 5 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err
done
python tools/bench_fir.py --taps 51 > $O/bench_fir_51.json 2> $O/bench_fir_51.err
python tools/bench_fir.py --taps 97 > $O/bench_fir_97.json 2> $O/bench_fir_97.err
python tools/bench_operator.py > $O/bench_operator_per_draw.json 2> $O/bench_operator.err
for tool in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool python tools/sanitize.py > $O/compute_sanitizer_$tool.txt 2>&1; echo "$tool rc=$?" >> $O/compute_sanitizer_$tool.txt
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(d["value"] / 1e6, 2), "Mframes/s", "frac", round(d["roofline"]["frac"], 3) if d.get("roofline") else None,
              "e2e", round(d["e2e"]["value"] / 1e6, 2) if d.get("e2e") else None)
    except Exception as e:
        print(f, "ERR", e)
PY
