"""CPU suite: the N > 1 path with world_size 2 over gloo.  The shard logic
(frame ranges, overlap spans, scatter of raw IQ, gather of rows, sweep
assignment) is what runs here; the per-rank transform is the oracle standing in
for the engine, because this box has no GPU (on the GPU box each rank calls
fsea_exec_* on its own device instead -- bench.py --gpus N)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import synth_u8
from frequensea_b200 import shard


def test_frame_shard_partition():
    for n_frames in (0, 1, 7, 128, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard.frame_shard(n_frames, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == n_frames
            pos = 0
            for first, count in spans:
                assert first == pos
                pos += count
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert shard.shard_samples(10, 4, 512, 1024) == (5120, 3 * 512 + 1024)
    owners = [f % 8 for f in range(1000)]
    for r in range(8):
        assert shard.sweep_shard(1000, r, 8) == [f for f in range(1000) if owners[f] == r]


def _worker(rank, world, port, n, hop, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import orc
    full = torch.from_numpy(synth_u8((n_frames - 1) * hop + n, seed=11)) if rank == 0 else None
    mine = shard.scatter_iq(full, n_frames, hop, n, rank, world)
    first, count = shard.frame_shard(n_frames, rank, world)
    rows = orc.engine_batch(mine.numpy(), n, hop, count, orc.IN_U8, orc.WIN_HANN, orc.OUT_LOGMAG)
    out = shard.gather_rows(torch.from_numpy(rows), n_frames, rank, world)
    if rank == 0:
        ref = orc.engine_batch(full.numpy(), n, hop, n_frames, orc.IN_U8, orc.WIN_HANN, orc.OUT_LOGMAG)
        q.put(bool(np.array_equal(out.numpy(), ref)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,hop,n_frames", [(256, 128, 13), (1024, 1024, 8)])
def test_two_rank_scatter_transform_gather(n, hop, n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200) + n // 256
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, hop, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def _sweep_worker(rank, world, port, n_freqs, n, rows, step, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from frequensea_b200 import sweep
    from conftest import OracleEngine
    stream = lambda f: sweep.synth_frequency(f, n, rows, quiet_every=4)
    per = (n_freqs + world - 1) // world
    res = sweep.sweep(OracleEngine(n, rows, per, gate_rows=8), n_freqs, n, rows, step, source=stream, rank=rank, world=world)
    # mode="stripes": nothing travels but the gate decisions; every rank keeps the images of its frequencies side by side
    st = sweep.sweep(OracleEngine(n, rows, per, gate_rows=8), n_freqs, n, rows, n, source=stream, rank=rank, world=world, mode="stripes")
    stripes = [None] * world
    dist.all_gather_object(stripes, (st["frequencies"], np.asarray(st["stripe"])))
    if rank == 0:
        one = sweep.sweep(OracleEngine(n, rows, n_freqs, gate_rows=8), n_freqs, n, rows, step, source=stream)   # the same job on one rank
        same = all(np.array_equal(np.asarray(res[k]), np.asarray(one[k])) for k in ("means", "keep", "images", "stitched"))
        same = same and np.array_equal(st["keep"], one["keep"]) and np.array_equal(st["means"], one["means"])
        for freqs, stripe in stripes:
            assert stripe.shape == (rows, per * n)
            for j, f in enumerate(freqs):
                same = same and np.array_equal(stripe[:, j * n:(j + 1) * n], np.asarray(one["images"][f]))
        q.put((bool(same), res["keep"].tolist(), tuple(res["stitched"].shape)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sweep_gate_and_stitch():
    """cfg4 workflow (c/fft-batch-broad.c gate + pixels, c/fft-stitch-broad.c): frequency f -> rank f mod 2,
    images and gate decisions gathered on rank 0 and stitched == the one-rank job; every 4th frequency
    is noise-only and must be gated off, the others kept."""
    n_freqs, n, rows, step = 7, 256, 12, 100
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29850 + (os.getpid() % 100)
    procs = [ctx.Process(target=_sweep_worker, args=(r, 2, port, n_freqs, n, rows, step, q)) for r in range(2)]
    for p in procs:
        p.start()
    same, keep, shape = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same
    assert keep == [f % 4 != 3 for f in range(n_freqs)]
    assert tuple(shape) == (rows, n + (n_freqs - 1) * step)
