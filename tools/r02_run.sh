O=gpurun_out/r02e; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; tail -5 $O/pytest_gpu.txt
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02e/bench.json").read().strip().splitlines()[-1])
    print("value", d["value"]/1e6, "frac", d["roofline"]["frac"], "burst_frac", d["roofline"]["burst_frac"], "clocks", d["clocks"], "bench_s", d["bench_seconds"])
    print("e2e", d["e2e"])
    for k, v in d["secondary"].items():
        if v is None: print(k, None); continue
        if "error" in v or "skipped" in v: print(k, v); continue
        if k == "sweep": print(k, json.dumps(v)[:1500]); continue
        if k == "sustained": print(k, json.dumps(v)[:800]); continue
        print(k, round(v["value"]/1e6, 2), "Mframes/s frac", round(v["roofline"]["frac"], 3), v["kernel"][:40], v["parity_gate"], v["clocks"])
except Exception as e:
    print("ERR", e)
PY
timeout 300 python tools/sweep_mgpu.py --gpus 1 --freqs 100 --rows 512 --mode gather > $O/mgpu_1gpu_gather.json 2> $O/mgpu.err; cat $O/mgpu_1gpu_gather.json; tail -3 $O/mgpu.err
timeout 300 python tools/sweep_mgpu.py --gpus 1 --freqs 100 --rows 512 --mode p2p > $O/mgpu_1gpu_p2p.json 2>> $O/mgpu.err; cat $O/mgpu_1gpu_p2p.json
