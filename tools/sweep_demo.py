#!/usr/bin/env python
"""BASELINE cfg4 as a job on N GPUs (one rank per GPU, torchrun): centre frequency f -> rank f mod N, a
device-resident engine per rank (fsea_sweep_*), u8 images gathered over NCCL and stitched on rank 0 -- or
stored straight into rank 0's canvas over NVLink (--mode p2p) -- and checked against the same job run by
rank 0 alone.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/sweep_demo.py [--freqs 64] [--rows 64] [--fft-size 16384] [--mode gather|p2p]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from frequensea_b200 import sweep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--freqs", type=int, default=64)
    ap.add_argument("--rows", type=int, default=64)
    ap.add_argument("--fft-size", dest="n", type=int, default=16384)
    ap.add_argument("--mode", default="gather", choices=["gather", "p2p"])
    ap.add_argument("--step", type=int, default=0)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    step = a.step or (a.n if a.mode == "p2p" else a.n // 4)            # WIDTH_STEP = N / (fs / step)
    per = (a.freqs + world - 1) // world
    eng = sweep.DeviceEngine(a.n, a.rows, per, device=local)
    sweep.sweep(eng, min(a.freqs, world), a.n, a.rows, step, rank=rank, world=world, mode=a.mode)        # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = sweep.sweep(eng, a.freqs, a.n, a.rows, step, rank=rank, world=world, mode=a.mode, keep_images=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        one = sweep.sweep(sweep.DeviceEngine(a.n, a.rows, a.freqs, device=local), a.freqs, a.n, a.rows, step, keep_images=False)
        differ = [k for k in ("means", "keep") if not np.array_equal(res[k], one[k])]
        if not torch.equal(res["stitched"], one["stitched"]):
            differ.append("stitched")
        print(json.dumps({"workload": "cfg4 sweep job (torchrun, one rank per GPU)", "n_gpus": world, "freqs": a.freqs, "rows": a.rows,
                          "fft_size": a.n, "mode": a.mode, "width_step": step, "kept": int(res["keep"].sum()),
                          "stitched_shape": list(res["stitched"].shape), "sharded_equals_single_rank": not differ, "differ": differ,
                          "seconds_incl_source": dt, "rows_per_s_incl_source": float(res["keep"].sum()) * a.rows / dt}))
        assert not differ
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
