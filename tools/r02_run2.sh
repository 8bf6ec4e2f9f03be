O=gpurun_out/r02g; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 300 python tools/sweep_mgpu.py --gpus 2 --freqs 40 --rows 64 --mode gather --check > $O/mgpu_2gpu_gather_check.json 2> $O/mgpu.err; cat $O/mgpu_2gpu_gather_check.json; tail -3 $O/mgpu.err
timeout 300 python tools/sweep_mgpu.py --gpus 2 --freqs 40 --rows 64 --mode p2p --check --f32-gather-mib 0 > $O/mgpu_2gpu_p2p_check.json 2>> $O/mgpu.err; cat $O/mgpu_2gpu_p2p_check.json
timeout 300 python tools/sweep_mgpu.py --gpus 2 --freqs 1000 --rows 512 --mode p2p --f32-gather-mib 0 > $O/mgpu_2gpu_p2p_full.json 2>> $O/mgpu.err; cat $O/mgpu_2gpu_p2p_full.json
timeout 300 python tools/sweep_mgpu.py --gpus 2 --freqs 1000 --rows 512 --mode gather > $O/mgpu_2gpu_gather_full.json 2>> $O/mgpu.err; cat $O/mgpu_2gpu_gather_full.json
timeout 300 python tools/sweep_mgpu.py --gpus 1 --freqs 1000 --rows 512 --mode p2p > $O/mgpu_1gpu_p2p_full.json 2>> $O/mgpu.err; cat $O/mgpu_1gpu_p2p_full.json
tail -3 $O/mgpu.err
for mode in gather p2p; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/sweep_demo.py --freqs 40 --rows 64 --mode $mode > $O/demo_2gpu_$mode.json 2> $O/demo_$mode.err; cat $O/demo_2gpu_$mode.json; tail -2 $O/demo_$mode.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; tail -3 $O/bench_2gpu.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02f/bench_2gpu.json").read().strip().splitlines()[-1])
    print("value", d["value"]/1e6, "frac", d["roofline"]["frac"], "burst_frac", d["roofline"]["burst_frac"], "bench_s", d["bench_seconds"])
    print("e2e", d["e2e"])
    for k, v in d["secondary"].items():
        print(k, json.dumps(v)[:1200])
except Exception as e:
    print("ERR", e)
PY
