O=gpurun_out/r02m; mkdir -p $O
G=$(nvidia-smi -L | wc -l); echo "gpus $G" > $O/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $G --steps 20 --warmup 5 > $O/bench_${G}gpu.json 2> $O/bench_${G}gpu.err; tail -2 $O/bench_${G}gpu.err | cut -c1-300
timeout 300 python tools/sweep_mgpu.py --gpus $G --freqs 1000 --rows 512 --mode p2p --f32-gather-mib 0 > $O/mgpu_${G}gpu_p2p_full.json 2> $O/mgpu.err; cat $O/mgpu_${G}gpu_p2p_full.json
timeout 300 python tools/sweep_mgpu.py --gpus $G --freqs 1000 --rows 512 --mode gather > $O/mgpu_${G}gpu_gather_full.json 2>> $O/mgpu.err; cat $O/mgpu_${G}gpu_gather_full.json
timeout 300 python tools/sweep_mgpu.py --gpus $G --freqs 200 --rows 128 --mode gather --check --f32-gather-mib 0 > $O/mgpu_${G}gpu_gather_check.json 2>> $O/mgpu.err; cat $O/mgpu_${G}gpu_gather_check.json
timeout 300 python tools/sweep_mgpu.py --gpus $G --freqs 200 --rows 128 --mode p2p --check --f32-gather-mib 0 > $O/mgpu_${G}gpu_p2p_check.json 2>> $O/mgpu.err; cat $O/mgpu_${G}gpu_p2p_check.json
tail -3 $O/mgpu.err
python - <<PY
import json
try:
    d = json.loads(open("$O/bench_${G}gpu.json").read().strip().splitlines()[-1])
    print("value", d["value"]/1e6, "frac", d["roofline"]["frac"], "burst_frac", d["roofline"]["burst_frac"], "bench_s", d["bench_seconds"], d["clocks"])
    print("e2e", json.dumps(d["e2e"])[:700])
    for k, v in d["secondary"].items():
        print(k, json.dumps(v)[:1400])
except Exception as e:
    print("ERR", e)
PY
