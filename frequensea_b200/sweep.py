"""The multi-centre-frequency sweep (BASELINE cfg4) as one sharded job.

What the reference does per centre frequency (c/fft-batch-broad.c:56-129,189-209): retune, take
rows of an N-point spectrum, after GATE_ROWS rows look at the mean magnitude and give the
frequency up if nothing is there (:80-98), otherwise finish HEIGHT rows of u8 pixels
(S = 5, DC patch, :108-118) -- one image per frequency; c/fft-stitch-broad.c:28-36,66-88 then
max-blends the images side by side, image i at column i * WIDTH_STEP.

Here centre frequency f belongs to rank f mod world; every rank gates and
transforms its own frequencies on its own GPU with no collective in the data path; the u8 images
(4x less traffic than float32 spectra) and the gate decisions are gathered on one rank, which
stitches them on its device.  `ops` is the per-rank arithmetic: EngineOps (the CUDA engine) in the
product, a stand-in only where a test says so.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import numpy as np

from . import shard

GATE_ROWS = 100          # c/fft-batch-broad.c:80
GATE_THRESHOLD = 1.1     # c/fft-batch-broad.c:93
PIX_SCALE = 5.0          # c/fft-batch-broad.c:112-113


def synth_frequency(freq_idx: int, n: int, rows: int, quiet_every: int = 10) -> np.ndarray:
    """SURVEY.md 8(d) cfg4 generator: per-frequency int8 IQ (two's complement, as bytes), seeded by the
    frequency index: noise + one tone whose bin depends on freq_idx (a mis-routed shard shows), and
    every `quiet_every`-th frequency noise-only at a level the gate rejects."""
    rng = np.random.default_rng(1000 + freq_idx)
    count = n * rows
    quiet = (freq_idx % quiet_every) == quiet_every - 1
    # gate mean = DC share 0.707 (uncentred samples, c/fft-batch-broad.c:65-68) + noise sigma sqrt(pi N / 2) / 256:
    # sigma = 0.4 stays under the 1.1 threshold up to N = 16384, sigma = 12 is over it from N = 64
    sigma = 0.4 if quiet else 12.0
    i = rng.normal(0, sigma, count)
    q = rng.normal(0, sigma, count)
    if not quiet:
        k = (37 * freq_idx + 11) % n
        ph = 2 * np.pi * (k - n // 2) / n * np.arange(count)
        i += 40 * np.cos(ph)
        q += 40 * np.sin(ph)
    out = np.empty(2 * count, dtype=np.int8)
    out[0::2] = np.clip(np.round(i), -128, 127)
    out[1::2] = np.clip(np.round(q), -128, 127)
    return out.view(np.uint8)


class EngineOps:
    """The engine on this rank's GPU: gate mean (fsea_gate_mean_host), pixel rows, stitch."""

    def __init__(self, n: int, device: int = 0, pix_scale: float = PIX_SCALE):
        from . import binding as B
        self._B = B
        self.device = device
        self.pix = B.Plan(n, in_format=B.IN_S8, out_mode=B.OUT_PIX, pix_scale=pix_scale, flags=B.FLAG_DC_PATCH, device=device)
        self.mag = B.Plan(n, in_format=B.IN_S8, out_mode=B.OUT_MAG, device=device)      # |X| incl. the DC bin, as :84-90 sums it

    def gate_mean(self, raw: np.ndarray, rows: int) -> float:
        return self.mag.gate_mean(raw, rows)

    def pixel_rows(self, raw: np.ndarray, rows: int) -> np.ndarray:
        return self.pix.exec_host(raw, rows)

    def stitch(self, images: np.ndarray, width_step: int) -> np.ndarray:
        return self._B.stitch(images, width_step, self.device)


def sweep_local(ops, freqs: List[int], stream: Callable[[int], np.ndarray], n: int, rows: int,
                gate_rows: int = GATE_ROWS, gate_threshold: float = GATE_THRESHOLD):
    """This rank's frequencies: (means, keep flags, images[len(freqs), rows, n] with zeros where gated off)."""
    means = np.zeros(len(freqs), dtype=np.float64)
    keep = np.zeros(len(freqs), dtype=bool)
    images = np.zeros((len(freqs), rows, n), dtype=np.uint8)
    for j, f in enumerate(freqs):
        raw = stream(f)
        means[j] = ops.gate_mean(raw, min(gate_rows, rows))
        keep[j] = not (means[j] < gate_threshold)                  # "interesting", c/fft-batch-broad.c:93-97
        if keep[j]:
            images[j] = ops.pixel_rows(raw, rows)
    return means, keep, images


def sweep(ops, n_freqs: int, stream: Callable[[int], np.ndarray], n: int, rows: int, width_step: int,
          rank: int = 0, world: int = 1, dst: int = 0, group=None,
          gate_rows: int = GATE_ROWS, gate_threshold: float = GATE_THRESHOLD) -> Optional[Dict[str, np.ndarray]]:
    """The whole sweep.  On `dst`: {"means", "keep", "images" (frequency order), "stitched"}; None elsewhere."""
    mine = shard.sweep_shard(n_freqs, rank, world)
    means, keep, images = sweep_local(ops, mine, stream, n, rows, gate_rows, gate_threshold)
    if world == 1:
        all_means, all_keep, all_images = means, keep, images
    else:
        import torch
        import torch.distributed as dist
        per = (n_freqs + world - 1) // world                         # pad every shard to the same length
        # NCCL moves device memory (u8 pixel images over NVLink); gloo, in the CPU tests, host memory
        dev = torch.device("cuda", getattr(ops, "device", 0)) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        pad_i = torch.zeros((per, rows, n), dtype=torch.uint8, device=dev)
        pad_m = torch.zeros(per, dtype=torch.float64, device=dev)
        pad_k = torch.zeros(per, dtype=torch.uint8, device=dev)
        pad_i[: len(mine)] = torch.from_numpy(images).to(dev)
        pad_m[: len(mine)] = torch.from_numpy(means).to(dev)
        pad_k[: len(mine)] = torch.from_numpy(keep.astype(np.uint8)).to(dev)
        if rank == dst:
            gi = [torch.empty_like(pad_i) for _ in range(world)]
            gm = [torch.empty_like(pad_m) for _ in range(world)]
            gk = [torch.empty_like(pad_k) for _ in range(world)]
        else:
            gi = gm = gk = None
        dist.gather(pad_i, gi, dst=dst, group=group)
        dist.gather(pad_m, gm, dst=dst, group=group)
        dist.gather(pad_k, gk, dst=dst, group=group)
        if rank != dst:
            return None
        all_images = np.zeros((n_freqs, rows, n), dtype=np.uint8)
        all_means = np.zeros(n_freqs, dtype=np.float64)
        all_keep = np.zeros(n_freqs, dtype=bool)
        for r in range(world):
            fr = shard.sweep_shard(n_freqs, r, world)
            all_images[fr] = gi[r][: len(fr)].cpu().numpy()
            all_means[fr] = gm[r][: len(fr)].cpu().numpy()
            all_keep[fr] = gk[r][: len(fr)].cpu().numpy().astype(bool)
    stitched = ops.stitch(all_images, width_step)
    return {"means": all_means, "keep": all_keep, "images": all_images, "stitched": stitched}
