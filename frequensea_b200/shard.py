"""Sharding of the IQ -> spectrum path over the GPUs of one box.

Frames are independent (the only per-stream state of the reference is the row
order of its history image, src/nrf.c:617), and so are the centre frequencies
of a sweep (c/fft-batch-broad.c:189-209): each rank transforms its own shard
with its own plan and NO collective sits inside the data path.  NCCL (or gloo
in the CPU tests) is used only to move whole shards: scatter raw IQ, gather
finished rows.

  streaming:  contiguous frame ranges; with overlap (hop < N) a shard also needs
              the N - hop samples that precede its successor's first frame
  sweep:      centre frequency f -> rank f mod world
"""
from __future__ import annotations

from typing import List, Tuple


def frame_shard(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """(first_frame, count) of `rank`: contiguous, sizes differ by at most one."""
    base, extra = divmod(n_frames, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def shard_samples(first: int, count: int, hop: int, n: int) -> Tuple[int, int]:
    """(first_sample, n_samples) a rank must hold to transform frames [first, first+count)."""
    if count <= 0:
        return first * hop, 0
    return first * hop, (count - 1) * hop + n


def sweep_shard(n_freqs: int, rank: int, world: int) -> List[int]:
    """Centre-frequency indices owned by `rank` (f mod world == rank)."""
    return list(range(rank, n_freqs, world))


def gather_rows(local_rows, n_frames: int, rank: int, world: int, dst: int = 0, group=None):
    """Collect every rank's finished rows on `dst` in frame order (torch.distributed).
    Returns the full (n_frames, N) tensor on dst, None elsewhere."""
    import torch
    import torch.distributed as dist
    counts = [frame_shard(n_frames, r, world)[1] for r in range(world)]
    width = local_rows.shape[1]
    pad = max(counts)
    buf = torch.zeros((pad, width), dtype=local_rows.dtype, device=local_rows.device)
    buf[: local_rows.shape[0]] = local_rows
    if rank == dst:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.gather(buf, parts, dst=dst, group=group)
        return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    dist.gather(buf, None, dst=dst, group=group)
    return None


def scatter_iq(full_iq, n_frames: int, hop: int, n: int, rank: int, world: int, src: int = 0, group=None):
    """Hand every rank the raw bytes of its frame shard (2 bytes per complex sample)."""
    import torch
    import torch.distributed as dist
    spans = [shard_samples(*frame_shard(n_frames, r, world), hop, n) for r in range(world)]
    mine = torch.empty(2 * spans[rank][1], dtype=torch.uint8, device=full_iq.device if full_iq is not None else "cpu")
    if rank == src:
        for r, (s0, ns) in enumerate(spans):
            piece = full_iq[2 * s0: 2 * (s0 + ns)]
            if r == src:
                mine.copy_(piece)
            elif ns:
                dist.send(piece.contiguous(), dst=r, group=group)
    elif spans[rank][1]:
        dist.recv(mine, src=src, group=group)
    return mine
