#!/usr/bin/env python
"""BASELINE cfg4 as a job on N GPUs (one rank per GPU, torchrun): centre frequency f -> rank f mod N,
per-rank gate + pixel rows on the rank's own device, u8 images gathered over NCCL, stitched on rank 0,
and checked against the same job run by rank 0 alone.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/sweep_demo.py [--freqs 64] [--rows 64] [--fft-size 16384]
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
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    step = a.n // 4                                                    # WIDTH_STEP = N / (fs / step): 20 MS/s, 5 MHz steps
    cache = {}

    def stream(f):
        if f not in cache:
            cache[f] = sweep.synth_frequency(f, a.n, a.rows)
        return cache[f]

    ops = sweep.EngineOps(a.n, device=local)
    for f in range(rank, a.freqs, world):
        stream(f)                                                      # generate outside the timed region
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = sweep.sweep(ops, a.freqs, stream, a.n, a.rows, step, rank=rank, world=world)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        one = sweep.sweep(ops, a.freqs, stream, a.n, a.rows, step)
        differ = [k for k in ("means", "keep", "images", "stitched") if not np.array_equal(res[k], one[k])]
        same = not differ
        print(json.dumps({"workload": "cfg4 sweep job", "n_gpus": world, "freqs": a.freqs, "rows": a.rows, "fft_size": a.n,
                          "kept": int(res["keep"].sum()), "stitched_shape": list(res["stitched"].shape),
                          "sharded_equals_single_rank": bool(same), "differ": differ, "seconds": dt,
                          "rows_per_s": float(res["keep"].sum()) * a.rows / dt}))
        assert same
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
