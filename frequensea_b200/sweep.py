"""The multi-centre-frequency sweep (BASELINE cfg4) as one sharded, device-resident job.

What the reference does per centre frequency (c/fft-batch-broad.c:56-129,189-209): retune, take
rows of an N-point spectrum, after GATE_ROWS rows look at the mean magnitude and give the
frequency up if nothing is there (:80-98), otherwise finish HEIGHT rows of u8 pixels
(S = 5, DC patch, :108-118) -- one image per frequency; c/fft-stitch-broad.c:28-36,66-88 then
max-blends the images side by side, image i at column i * WIDTH_STEP.

Here centre frequency f belongs to rank f mod world.  Every rank owns a
device-resident engine (fsea_sweep_*, include/fsea.h): the raw IQ of a frequency is uploaded (or
drawn on the device) once, the gate sum is reduced on the device, the keep flags cross to the host
in one small copy, and the pixel rows never leave device memory.  Getting them to the rank that
owns the stitched picture is the job's only exchange, in one of two ways:

  mode="gather"  NCCL gather of the u8 images (device -> device over NVLink; 4x less traffic than
                 float32 spectra), max-blend stitch on the destination's device;
  mode="p2p"     (WIDTH_STEP == N: the images tile the picture) the destination shares its canvas
                 over CUDA IPC and every rank's transform kernel stores its pixel rows straight
                 into it -- the kernel's stores ARE the gather (NVLink peer writes), there is no
                 collective and no stitch pass.
  mode="stripes" no exchange at all: every rank keeps the pixel rows of ITS frequencies as a stripe canvas
                 [rows, per N] in its own memory (frequency j world + rank at column j N) -- SURVEY.md 8(e)'s
                 "spectra left where they are produced"; a consumer reads the stripes in parallel (8 PCIe
                 links instead of one).  Every rank gets {"stripe": ...} back.

`engine` is DeviceEngine in the product; the CPU test-suite passes a stand-in with the same methods.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

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


class DeviceEngine:
    """This rank's GPU: fsea_sweep (raw -> gate -> pixel rows, device-resident) + stitch."""

    def __init__(self, n: int, rows: int, capacity: int, device: int = 0, gate_rows: int = GATE_ROWS,
                 gate_threshold: float = GATE_THRESHOLD, pix_scale: float = PIX_SCALE, pixel_arith: int = 0):
        import torch
        from . import binding as B
        self._B, self._torch = B, torch
        self.n, self.rows, self.capacity, self.device = n, rows, capacity, device
        self.tdev = torch.device("cuda", device)
        self.eng = B.Sweep(n, rows, capacity, gate_rows=gate_rows, gate_threshold=gate_threshold, pix_scale=pix_scale,
                           in_format=B.IN_S8, device=device, own_raw=True, own_images=False, pixel_arith=pixel_arith)

    def submit(self, slot: int, freq_idx: int, raw: Optional[np.ndarray] = None, quiet_every: int = 10) -> None:
        """raw IQ of a frequency: host bytes (uploaded once) or, with raw=None, the 8(d) source drawn on the device"""
        if raw is None:
            self.eng.submit_synth(slot, freq_idx, quiet_every)
        else:
            self.eng.submit_host(slot, raw)

    def gate(self, n_slots: int):
        return self.eng.gate(n_slots)

    def render_images(self, n_slots: int, per: int):
        """[per, rows, n] u8 on the device (slots >= n_slots zero): what a gather moves"""
        torch = self._torch
        img = torch.zeros((per, self.rows, self.n), dtype=torch.uint8, device=self.tdev)
        torch.cuda.current_stream(self.tdev).synchronize()          # the zero fill ran on torch's stream
        self.eng.render(n_slots, img.data_ptr(), self.n, 0, self.rows * self.n)
        self.eng.sync()
        return img

    def render_into(self, n_slots: int, ptr: int, stride: int, col0: int, col_step: int) -> None:
        self.eng.render(n_slots, ptr, stride, col0, col_step)
        self.eng.sync()

    def render_stripe(self, n_slots: int, per: int):
        """[rows, per N] u8 on the device: slot j's image at column j N (gated-off and unused slots zero)"""
        stripe = self.new_canvas(self.rows, per * self.n)
        self._torch.cuda.current_stream(self.tdev).synchronize()
        self.render_into(n_slots, stripe.data_ptr(), per * self.n, 0, self.n)
        return stripe

    def new_canvas(self, rows: int, width: int):
        return self._torch.zeros((rows, width), dtype=self._torch.uint8, device=self.tdev)

    def stitch(self, gathered, n_freqs: int, world: int, per: int, width_step: int):
        """gathered: [world * per, rows, n] u8 on this device, rank-major as dist.gather delivers it"""
        torch, B = self._torch, self._B
        out = torch.empty((self.rows, self.n + (n_freqs - 1) * width_step), dtype=torch.uint8, device=self.tdev)
        stream = torch.cuda.current_stream(self.tdev)
        B._check(B.lib().fsea_stitch_sharded_device(gathered.data_ptr(), n_freqs, world, per, self.rows, self.n, width_step,
                                                    out.data_ptr(), stream.cuda_stream), "fsea_stitch_sharded_device")
        stream.synchronize()
        return out

    @property
    def launches(self) -> int:
        return self.eng.launches


def _share_canvas(engine, rows: int, width: int, rank: int, dst: int, group):
    """The destination's canvas as a device pointer every rank of the box can store through (fsea_ipc_*: CUDA IPC,
    peer access over NVLink).  Returns (canvas tensor on dst or None, pointer, the buffer object to keep alive)."""
    import torch
    import torch.distributed as dist
    from .binding import SharedDeviceBuffer
    nbytes = rows * width
    buf, handle = None, [None]
    if rank == dst:
        buf = SharedDeviceBuffer.alloc(engine.device, nbytes)
        handle = [buf.handle]
    dist.broadcast_object_list(handle, src=dst, group=group)
    if rank == dst:
        canvas = torch.as_tensor(buf, device=engine.tdev).view(rows, width)     # zero-copy view of the shared allocation
        canvas.zero_()
        torch.cuda.synchronize(engine.tdev)
        dist.barrier(group=group)                                               # nobody stores before the canvas is clear
        return canvas, buf.ptr, buf
    buf = SharedDeviceBuffer.open(engine.device, handle[0], nbytes)
    dist.barrier(group=group)
    return None, buf.ptr, buf


def prepare_canvas(engine, n_freqs: int, n: int, rows: int, rank: int = 0, world: int = 1, dst: int = 0, group=None):
    """mode="p2p" across ranks: allocate the destination's canvas and share it ONCE (the allocation, the clear and the
    handle exchange are set-up, not part of a sweep); pass the result to sweep(canvas=...) as often as wanted."""
    if world == 1:
        cv = engine.new_canvas(rows, n * n_freqs)
        return cv, cv.data_ptr(), None
    return _share_canvas(engine, rows, n * n_freqs, rank, dst, group)


def sweep(engine, n_freqs: int, n: int, rows: int, width_step: int,
          source: Optional[Callable[[int], np.ndarray]] = None, rank: int = 0, world: int = 1, dst: int = 0, group=None,
          mode: str = "gather", quiet_every: int = 10, keep_images: bool = True, canvas=None) -> Optional[Dict]:
    """The whole sweep.  On `dst`: {"means", "keep" (numpy, frequency order), "stitched" (tensor on the engine's
    device, or numpy for a host stand-in), "images" (frequency order, when keep_images and mode == "gather")};
    None elsewhere.  source(f) -> the raw bytes of frequency f; None = the device-drawn 8(d) source."""
    import torch
    mine = shard.sweep_shard(n_freqs, rank, world)
    per = (n_freqs + world - 1) // world                             # every shard padded to the same length
    for j, f in enumerate(mine):
        engine.submit(j, f, None if source is None else source(f), quiet_every)
    means, keep = engine.gate(len(mine))
    out_width = n + (n_freqs - 1) * width_step
    pad_m = np.zeros(per, dtype=np.float64); pad_m[: len(mine)] = means
    pad_k = np.zeros(per, dtype=np.uint8); pad_k[: len(mine)] = keep

    if world == 1:
        all_means, all_keep = means, np.asarray(keep, dtype=bool)
        if mode == "p2p" and width_step == n:
            if canvas is None:
                cv = engine.new_canvas(rows, out_width)
                torch.cuda.synchronize()
            else:
                cv = canvas[0]                                     # prepare_canvas(): allocated and cleared once, kept between sweeps
            engine.render_into(len(mine), cv.data_ptr(), out_width, 0, n)
            return {"means": all_means, "keep": all_keep, "stitched": cv}
        img = engine.render_images(len(mine), per)
        res = {"means": all_means, "keep": all_keep, "stitched": engine.stitch(img, n_freqs, 1, per, width_step)}
        if keep_images:
            res["images"] = img[:n_freqs]
        return res

    import torch.distributed as dist
    nccl = dist.get_backend(group) == "nccl"
    dev = engine.tdev if nccl else torch.device("cpu")
    t_m = torch.from_numpy(pad_m).to(dev)
    t_k = torch.from_numpy(pad_k).to(dev)
    gm = [torch.empty_like(t_m) for _ in range(world)] if rank == dst else None
    gk = [torch.empty_like(t_k) for _ in range(world)] if rank == dst else None
    dist.gather(t_m, gm, dst=dst, group=group)
    dist.gather(t_k, gk, dst=dst, group=group)

    res = None
    if mode == "stripes":
        if canvas is None:
            stripe = engine.render_stripe(len(mine), per)                         # slot j at column j N of this rank's stripe
        else:                                                                     # a stripe canvas the caller keeps between sweeps
            stripe = canvas
            engine.render_into(len(mine), stripe.data_ptr(), per * n, 0, n)
        res = {"stripe": stripe, "frequencies": list(mine)}
        if rank != dst:
            return res
    elif mode == "p2p" and width_step == n and nccl:
        own = canvas is None
        cv, ptr, alive = _share_canvas(engine, rows, out_width, rank, dst, group) if own else canvas
        engine.render_into(len(mine), ptr, out_width, rank * n, world * n)        # frequency j * world + rank at column (j world + rank) N
        dist.barrier(group=group)                                                 # every rank's stores have landed (render_into synchronised its stream)
        if rank == dst:
            res = {"stitched": cv, "_shared": alive}                              # (the picture lives in the shared allocation: keep it)
        elif own:
            alive.close()
    else:
        img = engine.render_images(len(mine), per)
        if rank == dst:
            gathered = torch.empty((world * per, rows, n), dtype=torch.uint8, device=img.device)
            dist.gather(img, list(gathered.view(world, per, rows, n).unbind(0)), dst=dst, group=group)
            res = {"stitched": engine.stitch(gathered, n_freqs, world, per, width_step)}
            if keep_images:
                order = [(f % world) * per + f // world for f in range(n_freqs)]
                res["images"] = gathered[torch.as_tensor(order, device=gathered.device)]
        else:
            dist.gather(img, None, dst=dst, group=group)
    if rank != dst:
        return None
    all_means = np.zeros(n_freqs, dtype=np.float64)
    all_keep = np.zeros(n_freqs, dtype=bool)
    for r in range(world):
        fr = shard.sweep_shard(n_freqs, r, world)
        all_means[fr] = gm[r][: len(fr)].cpu().numpy()
        all_keep[fr] = gk[r][: len(fr)].cpu().numpy().astype(bool)
    res["means"], res["keep"] = all_means, all_keep
    return res
