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


class OracleOps:
    """Stand-in for sweep.EngineOps in this CPU suite only (no GPU here): the oracle's arithmetic."""
    device = 0

    def __init__(self, n):
        self.n = n

    def gate_mean(self, raw, rows):
        from oracle import orc
        return float(orc.engine_batch(raw, self.n, self.n, rows, orc.IN_S8, orc.WIN_NONE, orc.OUT_MAG).astype(np.float64).mean())

    def pixel_rows(self, raw, rows):
        from oracle import orc
        return orc.engine_batch(raw, self.n, self.n, rows, orc.IN_S8, orc.WIN_NONE, orc.OUT_PIX, 5.0, flags=orc.FLAG_DC_PATCH)

    def stitch(self, images, step):
        from oracle import orc
        return orc.stitch(images, step)


def _sweep_worker(rank, world, port, n_freqs, n, rows, step, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from frequensea_b200 import sweep
    stream = lambda f: sweep.synth_frequency(f, n, rows, quiet_every=4)
    res = sweep.sweep(OracleOps(n), n_freqs, stream, n, rows, step, rank=rank, world=world, gate_rows=8)
    if rank == 0:
        one = sweep.sweep(OracleOps(n), n_freqs, stream, n, rows, step, gate_rows=8)      # the same job on one rank
        same = all(np.array_equal(res[k], one[k]) for k in ("means", "keep", "images", "stitched"))
        q.put((bool(same), res["keep"].tolist(), res["stitched"].shape))
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
