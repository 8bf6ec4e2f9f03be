# 8-GPU check at HEAD: the driver's scaling command (full bench line under torchrun) and the one-process sweep driver
O=gpurun_out/r02n; mkdir -p $O
G=$(nvidia-smi -L | wc -l); echo "gpus $G" > $O/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $G --steps 20 --warmup 5 > $O/bench_${G}gpu.json 2> $O/bench_${G}gpu.err; tail -2 $O/bench_${G}gpu.err | cut -c1-300
timeout 300 python tools/sweep_mgpu.py --gpus $G --freqs 1000 --rows 512 --mode p2p --f32-gather-mib 0 > $O/mgpu_${G}gpu_p2p_full.json 2> $O/mgpu.err; cat $O/mgpu_${G}gpu_p2p_full.json | cut -c1-800
timeout 300 python tools/sweep_mgpu.py --gpus $G --freqs 200 --rows 128 --mode p2p --check --f32-gather-mib 0 > $O/mgpu_${G}gpu_p2p_check.json 2>> $O/mgpu.err; cat $O/mgpu_${G}gpu_p2p_check.json | cut -c1-600
tail -3 $O/mgpu.err
python - <<PY
import json
try:
    d = json.loads(open("$O/bench_${G}gpu.json").read().strip().splitlines()[-1])
    print("value", d["value"]/1e6, "frac", d["roofline"]["frac"], "bench_s", d["bench_seconds"], d["clocks"])
    print("e2e", json.dumps(d["e2e"])[:500])
    for k, v in d["secondary"].items():
        if k == "sweep": print(k, json.dumps(v["modes"])[:2500])
        else: print(k, json.dumps({kk: vv for kk, vv in v.items() if kk in ("value", "ms_per_step", "roofline_frac", "error", "skipped")}))
except Exception as e:
    print("ERR", e)
PY
