#!/usr/bin/env python3
"""bench.py -- benchmark of the IQ -> spectrum hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload (BASELINE.json configs[1], cfg2): the nrf_fft path -- 1024-point FFT of every
frame of rfdata/rf-433.000.raw (two's-complement int8 IQ, 128 frames per 262144-byte block),
float32 magnitude with the reference's DC patch -- the capture block replayed to 2^18 frames
(512 MiB in, 1 GiB out: far larger than the 126 MB L2).  A STEP is R passes of the hot path over
that batch, R chosen so that the K timed steps last >= ~1.2 s (the clock sampler then sees the
kernel under sustained load); `burst` is the same kernel timed alone over K single launches, as in
round 1.  Metric: frames/s (whole job).

  value       device-resident (fsea_exec_device), CUDA events on the launching stream, max over ranks
  burst       K single launches (7 ms): the kernel at boost clocks
  e2e         the same batch through the C ABI with HOST buffers (fsea_exec_host): pinned host ->
              H2D -> kernel -> D2H -> pinned host; next to it the same copies WITHOUT the kernel
              (`copy_only`: what the box's PCIe links allow) and the u8-pixel form (3N bytes/frame)
  roofline    algorithmic bytes (2*hop + 4*N per frame) / kernel time against the
              measured HBM copy bandwidth (MEASURED_PEAKS.json)
  secondary   every other BASELINE configuration and the size ladder, measured the same way:
              cfg3 (8192 Hann 50 % overlap, the 8(d) tone source), cfg4 (16384-pt sweep rows -> u8
              pixels), cfg5 (65536), n128 .. n65536; `sustained` = cfg5 over a resident ring of
              4096 frames for >= 5 s; `sweep` = the cfg4 JOB (1000 centre frequencies x 512 rows,
              gate + pixels + stitched picture, device-resident engines, f -> rank f mod N)
  cpu_baseline  the reference's own nrf_fft_* (oracle/_ref) or the oracle port, timed on this box's
              host cores on a bounded sample of the same frames

Multi-GPU (torchrun, one rank per GPU): frames are independent, so each rank transforms its own
shard -- no data-path collective; weak scaling.  The sweep job's only exchange is
the u8 pictures (NCCL gather, or peer stores straight into the destination's canvas).
"""
import argparse
import ctypes as C
import datetime
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N = 1024
HOP = 1024
FRAMES_PER_BLOCK = 128
BLOCK = 262144
DEFAULT_FRAMES = 1 << 18          # per GPU per launch: 512 MiB in, 1 GiB out
METRIC = "fft_frames_per_s"
UNIT = "frames/s"


def _size_workload(n):
    frames = max(256, (1 << 28) // n)            # 512 MiB of raw IQ per launch
    return dict(n=n, hop=n, fmt="s8", window=False, out="mag", patch=True, frames=frames, source="uniform",
                desc=f"size ladder: {n}-pt FFT, int8 IQ -> f32 magnitude + DC patch")


WORKLOADS = {
    "cfg2": dict(n=1024, hop=1024, fmt="s8", window=False, out="mag", patch=True, frames=1 << 18, source="uniform",
                 desc="cfg2 nrf_fft path: 1024-pt FFT, all frames of rf-433.000.raw, int8 IQ -> f32 magnitude + DC patch"),
    "cfg3": dict(n=8192, hop=4096, fmt="u8", window=True, out="logmag", patch=False, frames=65535, source="cfg3",
                 desc="cfg3: 8192-pt FFT, Hann, 50% overlap, synthetic 2.4 MS/s uint8 IQ (3 tones + noise, SURVEY 8(d)) -> f32 log-magnitude"),
    "cfg4": dict(n=16384, hop=16384, fmt="s8", window=False, out="pix", patch=True, frames=1 << 14, scale=5.0, source="uniform",
                 desc="cfg4 sweep rows: 16384-pt FFT, int8 IQ -> u8 pixels (S=5, DC patch)"),
    "cfg5": dict(n=65536, hop=65536, fmt="s8", window=False, out="logmag", patch=False, frames=4096, source="cfg5",
                 desc="cfg5: 65536-pt FFT, synthetic 20 MS/s int8 HackRF IQ (noise + chirp, SURVEY 8(d)) -> f32 log-magnitude"),
}
for _n in (128, 256, 512, 2048, 4096, 8192, 16384, 32768, 65536):
    WORKLOADS[f"n{_n}"] = _size_workload(_n)
SECONDARY_ORDER = ["cfg3", "cfg4", "cfg5", "n2048", "n4096", "n8192", "n16384", "n32768", "n65536", "n512", "n256", "n128"]


def load_capture():
    z = np.load(os.path.join(ROOT, "tests", "golden", "iq_fixtures.npz"))
    return z["rf433"]


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md): one sampler runs
    for the whole bench, every timed region reads back the samples that fall inside it."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0, period_ms=50):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            import shutil
            cmd = ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", str(period_ms), "-i", str(gpu_index)]
            if shutil.which("stdbuf"):
                cmd = ["stdbuf", "-oL"] + cmd                  # line-buffered: a region reads its samples while the sampler runs
            self.proc = subprocess.Popen(cmd, stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    @staticmethod
    def now():
        return datetime.datetime.now()

    def region(self, t0, t1):
        """summary of the samples taken between the datetimes t0 and t1 (a little slack on both sides)"""
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.06)
        self.tmp.flush()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        slack = datetime.timedelta(milliseconds=20)
        for line in open(self.tmp.name):
            r = [x.strip() for x in line.strip().split(",")]
            if len(r) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f")
                if ts < t0 - slack or ts > t1 + slack:
                    continue
                sm.append(float(r[2])); mx.append(float(r[3])); power.append(float(r[4]))
                for nm, v in zip(names, r[6:10]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            hot = [s for s, p in zip(sm, power) if p >= 0.6 * max(power)] or sm
            out.update(sm_mhz=float(np.median(hot)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(power)), power_w_median=float(np.median(power)))
        return out

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            try:
                os.unlink(self.tmp.name)
            except OSError:
                pass
            self.proc = None


def dist_setup(gpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    return rank, world, local


def cpu_reference_rate(raw, threads, target_s=12.0):
    """Reference arm / CPU baseline: the reference's own nrf_fft_* (oracle/_ref) when
    it was compiled, else the oracle port, on a bounded sample of the workload."""
    from oracle import orc
    kind = "reference" if orc.have_ref() else "port"
    block = raw[:BLOCK]

    def run(frames, th):
        data = np.tile(block, (frames + FRAMES_PER_BLOCK - 1) // FRAMES_PER_BLOCK)
        out = np.empty((frames, N), dtype=np.float32)
        t0 = time.perf_counter()
        if kind == "reference":
            R = orc.ref()
            R.ref_run_frames.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_long, C.c_int, C.c_void_p, C.c_int]
            R.ref_run_frames.restype = None
            R.ref_run_frames(data.ctypes.data, N, HOP, frames, 1, out.ctypes.data, th)
        else:
            orc.engine_batch(data, N, HOP, frames, orc.IN_S8, orc.WIN_NONE, orc.OUT_MAG, flags=orc.FLAG_DC_PATCH, threads=th)
        return time.perf_counter() - t0

    probe = 4096
    dt = run(probe, threads)
    frames = int(max(probe, min(1 << 22, probe * target_s / max(dt, 1e-6))))
    frames -= frames % FRAMES_PER_BLOCK
    return kind, frames, run


# ----------------------------------------------------------------------------------------- helpers
class Ctx:
    """what every leg needs: rank / world, device, stream, barrier, the clock sampler"""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.args = args
        self.rank, self.world, self.local = dist_setup(args.gpus)
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.stream = torch.cuda.current_stream()
        self.clocks = ClockSampler(self.local) if self.rank == 0 else None
        self.t_start = time.perf_counter()

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def elapsed(self):
        return time.perf_counter() - self.t_start


def oracle_check(w, d_in, d_out, frames, frame_ids):
    """parity gate on the data that is timed: oracle on a few frames.  Returns (worst, ok, what)."""
    from oracle import orc
    n, hop = w["n"], w["hop"]
    fmt = orc.IN_S8 if w["fmt"] == "s8" else orc.IN_U8
    win = orc.WIN_HANN if w["window"] else orc.WIN_NONE
    out = {"mag": orc.OUT_MAG, "logmag": orc.OUT_LOGMAG, "pix": orc.OUT_PIX}[w["out"]]
    worst, strict, ndiff, total = 0.0, 0.0, 0, 0
    for f in frame_ids:
        raw = d_in[f * hop * 2:f * hop * 2 + 2 * n].cpu().numpy()
        ref = orc.engine_frame(raw, n, fmt, win, out, w.get("scale", 10.0), orc.FLAG_DC_PATCH if w["patch"] else 0).astype(np.float64)
        got = d_out[f].cpu().numpy().astype(np.float64)
        if w["out"] == "mag":
            rms = np.sqrt(np.mean(ref ** 2))
            worst = max(worst, float(np.max(np.abs(got - ref) / (rms + 0.1 * np.abs(ref)))))     # |d| <= 1e-5 rms + 1e-6 |ref|
            strict = max(strict, float(np.max(np.abs(got - ref)) / rms))                          # max|d| / rms
        elif w["out"] == "logmag":
            lin = 10 ** (ref / 20)
            m = lin >= 1e-3 * np.sqrt(np.mean(lin ** 2))
            worst = max(worst, float(np.max(np.abs(got - ref)[m])))
        else:
            d = np.abs(got - ref)
            worst = max(worst, float(d.max()))
            ndiff += int((d > 0).sum()); total += d.size
    if w["out"] == "mag":
        return {"max_err_over_rms_plus_0.1ref": worst, "strict_max_err_over_rms": strict, "bound": 1e-5, "frames": len(frame_ids)}, worst < 1e-5
    if w["out"] == "logmag":
        return {"max_db_err_bins_above_1e-3_rms": worst, "bound_db": 2e-3, "frames": len(frame_ids)}, worst < 2e-3
    return {"max_pixel_diff": worst, "pixels_differing": ndiff / max(total, 1), "bound": "<= 1 LSB", "frames": len(frame_ids)}, worst <= 1


def make_plan(fb, w, local, arith=0):
    fmt = fb.IN_S8 if w["fmt"] == "s8" else fb.IN_U8
    out_mode = {"mag": fb.OUT_MAG, "logmag": fb.OUT_LOGMAG, "pix": fb.OUT_PIX}[w["out"]]
    return fb.Plan(w["n"], hop=w["hop"], in_format=fmt, window=fb.WIN_HANN if w["window"] else fb.WIN_NONE, out_mode=out_mode,
                   flags=fb.FLAG_DC_PATCH if w["patch"] else 0, pix_scale=w.get("scale", 10.0), arith=arith, device=local)


def make_input(ctx, w, plan, frames):
    torch = ctx.torch
    nbytes = plan.input_bytes(frames)
    if w["source"] == "cfg3":
        from frequensea_b200 import synth
        return synth.synth_u8_cfg3(nbytes // 2, ctx.dev, seed=20260925 + ctx.rank), "SURVEY 8(d) cfg3 source drawn on the device (3 tones + N(0, 20^2), offset-binary uint8)"
    if w["source"] == "cfg5":
        from frequensea_b200 import synth
        return synth.synth_s8_cfg5(nbytes // 2, ctx.dev, seed=7 + ctx.rank), "SURVEY 8(d) cfg5 source drawn on the device (sigma-24 noise + chirp, int8)"
    g = torch.Generator(device=ctx.dev).manual_seed(20260925 + ctx.rank)
    return torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device=ctx.dev, generator=g), "uniform random bytes drawn on the device"


def timed_launches(ctx, plan, d_in, d_out, frames, steps, inner):
    """`steps` timed steps of `inner` launches each; CUDA events on the launching stream around every step.
    -> (total ms between the outer events, [ms per step], wall datetimes of the region)"""
    torch = ctx.torch
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = ctx.stream.cuda_stream
    ip, op = d_in.data_ptr(), d_out.data_ptr()
    ctx.barrier()
    t0 = ClockSampler.now()
    e0.record(ctx.stream)
    for a, b in ev:
        a.record(ctx.stream)
        for _ in range(inner):
            plan.exec_device(ip, op, frames, s)
        b.record(ctx.stream)
    e1.record(ctx.stream)
    ctx.barrier()
    t1 = ClockSampler.now()
    return e0.elapsed_time(e1), [a.elapsed_time(b) for a, b in ev], (t0, t1)


def measure_workload(ctx, fb, name, frames_override=None, min_region_s=0.35, arith=0):
    """one secondary entry: device-resident throughput + roofline of workload `name`, parity-gated"""
    torch = ctx.torch
    w = WORKLOADS[name]
    n, hop = w["n"], w["hop"]
    frames = frames_override or w["frames"]
    plan = make_plan(fb, w, ctx.local, arith)
    d_in, data_desc = make_input(ctx, w, plan, frames)
    pix = w["out"] == "pix"
    d_out = torch.empty((frames, n), dtype=torch.uint8 if pix else torch.float32, device=ctx.dev)
    s = ctx.stream.cuda_stream
    plan.exec_device(d_in.data_ptr(), d_out.data_ptr(), frames, s)
    torch.cuda.synchronize()
    gate, ok = oracle_check(w, d_in, d_out, frames, [0, 1, frames // 2, frames - 1])
    if not ok:
        raise RuntimeError(f"parity gate failed on {name}: {gate}")
    for _ in range(max(ctx.args.warmup, 3)):
        plan.exec_device(d_in.data_ptr(), d_out.data_ptr(), frames, s)
    torch.cuda.synchronize()
    # size the steps so that the timed region lasts >= min_region_s
    burst_ms, probe, _ = timed_launches(ctx, plan, d_in, d_out, frames, 5, 1)     # also the burst figure: 5 launches from a cool start
    burst_ms = ctx.max_over_ranks(burst_ms / 5)[0]
    steps = max(ctx.args.steps, 5)
    inner = max(1, int(np.ceil(min_region_s * 1e3 / (steps * max(min(probe), 1e-3)))))
    l0 = plan.launches
    total_ms, per_step, (t0, t1) = timed_launches(ctx, plan, d_in, d_out, frames, steps, inner)
    launches = plan.launches - l0
    ms_launch = ctx.max_over_ranks(total_ms / (steps * inner))[0]
    peak, peak_src = peak_hbm()
    alg = frames * (2 * hop + (1 if pix else 4) * n)
    achieved = alg / (ms_launch * 1e-3) / 1e9
    entry = {
        "workload": w["desc"], "value": ctx.world * frames / (ms_launch * 1e-3), "unit": UNIT, "ms_per_step": ms_launch * inner,
        "ms_per_launch": ms_launch, "launches_per_step": inner, "steps": steps, "frames_per_gpu_per_launch": frames,
        "iq_gsps": ctx.world * frames * hop / (ms_launch * 1e-3) / 1e9, "kernel": plan.kernel_name, "parity_gate": gate,
        "gpu_launches": launches, "data": data_desc,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                     "burst_frac": alg / (burst_ms * 1e-3) / 1e9 / peak, "burst_ms_per_launch": burst_ms,
                     "note": "frac: the >= 0.35 s region (power cap may apply, see clocks); burst_frac: 5 launches from a cool start"},
        "clocks": ctx.clocks.region(t0, t1) if ctx.clocks else None,
    }
    del d_in, d_out, plan
    torch.cuda.empty_cache()
    return entry


def sustained_cfg5(ctx, fb, seconds=5.0):
    """a resident ring of 4096 frames of 65536 points, transformed over and over for >= 5 s
    after warm-up: steady-state frames/s, GS/s, GB/s and x real time against 20 MS/s (305 frames/s)."""
    torch = ctx.torch
    w = WORKLOADS["cfg5"]
    n, frames = w["n"], 4096
    plan = make_plan(fb, w, ctx.local)
    d_in, data_desc = make_input(ctx, w, plan, frames)
    d_out = torch.empty((frames, n), dtype=torch.float32, device=ctx.dev)
    s = ctx.stream.cuda_stream
    for _ in range(3):
        plan.exec_device(d_in.data_ptr(), d_out.data_ptr(), frames, s)
    torch.cuda.synchronize()
    gate, ok = oracle_check(w, d_in, d_out, frames, [0, frames - 1])
    if not ok:
        raise RuntimeError(f"parity gate failed on sustained cfg5: {gate}")
    _, probe, _ = timed_launches(ctx, plan, d_in, d_out, frames, 2, 1)
    passes = max(8, int(np.ceil(seconds * 1e3 / max(min(probe), 1e-3))))
    steps = 8
    inner = (passes + steps - 1) // steps
    total_ms, per_step, (t0, t1) = timed_launches(ctx, plan, d_in, d_out, frames, steps, inner)
    total_ms = ctx.max_over_ranks(total_ms)[0]
    fps = ctx.world * frames * steps * inner / (total_ms * 1e-3)
    peak, _ = peak_hbm()
    gbs = fps / ctx.world * 6 * n / 1e9
    out = {"workload": "cfg5 sustained: resident ring of 4096 frames x 65536 points per GPU, int8 -> f32 log-magnitude",
           "seconds": total_ms * 1e-3, "passes_over_ring": steps * inner, "value": fps, "unit": UNIT, "iq_gsps": fps * n / 1e9,
           "gbs_per_gpu": gbs, "roofline_frac": gbs / peak, "x_real_time_at_20MSps": fps * n / 20e6,
           "first_vs_last_step_ms": [per_step[0], per_step[-1]], "kernel": plan.kernel_name, "parity_gate": gate, "data": data_desc,
           "clocks": ctx.clocks.region(t0, t1) if ctx.clocks else None}
    del d_in, d_out, plan
    torch.cuda.empty_cache()
    return out


def file_ingest(ctx, fb, frames=768, repeats=4):
    """a capture FILE (raw interleaved int8 as c/reader.c / c/rfcap.c write it) through
    fsea_exec_file -- pread() into pinned slots, H2D, the cfg5 transform, D2H -- against the 20 MS/s a HackRF delivers."""
    import tempfile
    torch = ctx.torch
    w = WORKLOADS["cfg5"]
    n = w["n"]
    rng = np.random.default_rng(11)
    raw = rng.integers(0, 256, size=2 * n * frames, dtype=np.uint8)
    d = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    with tempfile.NamedTemporaryFile(prefix="fsea_capture_", suffix=".raw", dir=d, delete=False) as fh:
        fh.write(raw.tobytes())
        path = fh.name
    try:
        plan = make_plan(fb, w, ctx.local)
        out = torch.empty((frames, n), dtype=torch.float32, pin_memory=True).numpy()
        plan.exec_file(path, frames, out=out)                       # warm: page cache, staging slots
        ref = plan.exec_host(raw, frames)
        same = bool(np.array_equal(out, ref))                       # the same bytes through the same kernel
        l0 = plan.launches
        t0 = time.perf_counter()
        for _ in range(repeats):
            plan.exec_file(path, frames, out=out)
        dt = (time.perf_counter() - t0) / repeats
        return {"workload": f"capture file of {frames} x {n}-pt frames ({raw.nbytes >> 20} MiB, int8 IQ) -> fsea_exec_file -> f32 log-magnitude rows on the host",
                "value": frames / dt, "unit": UNIT, "iq_gsps": frames * n / dt / 1e9, "x_real_time_at_20MSps": frames * n / dt / 20e6,
                "file_gbs": raw.nbytes / dt / 1e9, "equals_exec_host": same, "gpu_launches": plan.launches - l0,
                "where": "tmpfs" if d else "temp dir (page cache)"}
    finally:
        os.unlink(path)


def sweep_job(ctx, fb, n_freqs=1000, rows=512, n=16384):
    """The cfg4 JOB (c/fft-batch-broad.c:56-129,189-209 + c/fft-stitch-broad.c:66-88): 1000 centre frequencies x 512
    rows x 16384 points, frequency f -> rank f mod N, raw IQ resident per rank (drawn on the device, untimed), then
    timed: gate rows + decisions, pixel rows of the kept frequencies, the stitched picture on rank 0."""
    torch = ctx.torch
    from frequensea_b200 import sweep
    per = (n_freqs + ctx.world - 1) // ctx.world
    eng = sweep.DeviceEngine(n, rows, per, device=ctx.local)
    mine = list(range(ctx.rank, n_freqs, ctx.world))
    out = {"workload": f"cfg4 sweep job: {n_freqs} centre frequencies x {rows} rows x {n}-pt, int8 -> gate + u8 pixels (S=5, DC patch) + stitched picture",
           "n_freqs": n_freqs, "rows": rows, "fft_size": n, "partition": f"frequency f -> rank f mod {ctx.world}",
           "data": "SURVEY 8(d) cfg4 source drawn on the device per frequency (noise + tone at bin (37 f + 11) mod N; every 10th frequency quiet)"}
    # parity sample: a few of this rank's frequencies against the oracle (rank 0 only)
    for j, f in enumerate(mine):
        eng.submit(j, f)
    eng.eng.sync()
    modes = {}
    resident = _Resident(eng, len(mine))
    def column_checksums(pic, count):
        """one int64 per frequency: position-weighted sum of every 64th row of its N columns"""
        v = pic[::64].reshape(-1, count, n).to(torch.int64)
        w = torch.arange(1, n + 1, dtype=torch.int64, device=pic.device)
        return (v * w).sum(dim=(0, 2))
    full_sums = None
    for mode in ("gather", "p2p") + (("stripes",) if ctx.world > 1 else ()):
        try:
            canvas = sweep.prepare_canvas(eng, n_freqs, n, rows, ctx.rank, ctx.world) if mode == "p2p" else None
            if mode == "stripes":                                 # like the shared canvas: allocated and cleared once, outside the job
                canvas = eng.new_canvas(rows, per * n)
                torch.cuda.synchronize()
            # set-up run (connections, allocator), then the timed one
            res = sweep.sweep(resident, n_freqs, n, rows, n, rank=ctx.rank, world=ctx.world, mode=mode, keep_images=False, canvas=canvas)
            del res
            ctx.barrier()
            l0 = eng.launches
            t0 = time.perf_counter()
            res = sweep.sweep(resident, n_freqs, n, rows, n, rank=ctx.rank, world=ctx.world, mode=mode, keep_images=False, canvas=canvas)
            torch.cuda.synchronize()
            dt = ctx.max_over_ranks(time.perf_counter() - t0)[0]
            launches = eng.launches - l0
            if mode == "stripes":
                import torch.distributed as dist
                mine_sums = column_checksums(res["stripe"], per)
                all_sums = [torch.empty_like(mine_sums) for _ in range(ctx.world)]
                dist.all_gather(all_sums, mine_sums)
                if ctx.rank == 0:
                    kept = int(res["keep"].sum())
                    got = torch.stack(all_sums, dim=1).reshape(-1)[:n_freqs]          # frequency f = j world + r
                    modes[mode] = {"seconds": dt, "rows_per_s": kept * rows / dt, "kept": kept, "gated_off": n_freqs - kept,
                                   "stripe_shape": list(res["stripe"].shape), "gpu_launches_rank0": launches,
                                   "exchange": "none: every rank keeps the pixel rows of its frequencies (SURVEY 8(e) (i): results stay where they are produced)",
                                   "same_pixels_as_p2p": bool(full_sums is not None and torch.equal(got.cpu(), full_sums.cpu()))}
                del res
                canvas = None
                torch.cuda.empty_cache()
                continue
            if ctx.rank == 0:
                kept = int(res["keep"].sum())
                import hashlib
                pic = res["stitched"]
                if mode == "p2p" and ctx.world > 1:
                    full_sums = column_checksums(pic, n_freqs)
                # a cheap fingerprint of the 8 GB picture: column sums of every 64th row
                fp = hashlib.sha256(pic[::64].to(torch.int64).sum(dim=0).cpu().numpy().tobytes()).hexdigest()[:16]
                modes[mode] = {"seconds": dt, "rows_per_s": kept * rows / dt, "kept": kept, "gated_off": n_freqs - kept,
                               "picture_shape": list(pic.shape), "picture_fingerprint": fp, "gpu_launches_rank0": launches,
                               "exchange": ("NCCL gather of u8 images to rank 0 + max-blend stitch there" if mode == "gather" else
                                            "none: every rank's transform kernel stores its pixel rows into rank 0's canvas (CUDA IPC, NVLink peer stores)")
                                           if ctx.world > 1 else "one GPU: no exchange"}
                if mode == "p2p" and "gather" in modes and "picture_fingerprint" in modes["gather"]:
                    modes[mode]["same_picture_as_gather"] = modes["gather"]["picture_fingerprint"] == fp
                if "parity" not in out:
                    out["parity"] = _sweep_parity(eng, res, mine, n, rows)
            del res
            if canvas is not None and mode == "p2p" and canvas[2] is not None:
                canvas[2].close()
            canvas = None
            torch.cuda.empty_cache()
        except Exception as e:                                    # never lose the bench line over this leg
            modes[mode] = {"error": str(e)[:300]}
    out["modes"] = modes
    # the job's value is a COMPLETE picture on one GPU; the stripes leg is reported beside it
    best = max((m for k, m in modes.items() if "rows_per_s" in m and k != "stripes"), key=lambda m: m["rows_per_s"], default=None)
    if best:
        out["value"] = best["rows_per_s"]
        out["unit"] = "rows/s"
    del eng
    torch.cuda.empty_cache()
    return out


class _Resident:
    """sweep.sweep() over an engine whose raw IQ is already resident (bench: the source is drawn untimed)"""

    def __init__(self, eng, n_slots):
        self._e = eng
        self.tdev = eng.tdev

    def submit(self, *a, **k):
        pass

    def __getattr__(self, name):
        return getattr(self._e, name)


def _sweep_parity(eng, res, mine, n, rows):
    """rank 0: gate mean and pixel rows of two of its frequencies against the oracle on the bytes the device drew"""
    from oracle import orc
    out = {}
    try:
        worst_mean, worst_pix, ndiff, total = 0.0, 0, 0, 0
        for j in (0, min(9, len(mine) - 1)):
            f = mine[j]
            raw = eng.eng.read_raw(j)
            gr = min(100, rows)
            ref_mean = float(orc.engine_batch(raw, n, n, gr, orc.IN_S8, orc.WIN_NONE, orc.OUT_MAG).astype(np.float64).mean())
            worst_mean = max(worst_mean, abs(float(res["means"][f]) - ref_mean) / ref_mean)
            ref_keep = not (ref_mean < 1.1)
            assert bool(res["keep"][f]) == ref_keep
            sample_rows = [0, 1, rows // 2, rows - 1]
            got = res["stitched"][sample_rows, f * n:(f + 1) * n].cpu().numpy().astype(int)
            for k, r in enumerate(sample_rows):
                ref = orc.engine_frame(raw[r * 2 * n:], n, orc.IN_S8, orc.WIN_NONE, orc.OUT_PIX, 5.0, orc.FLAG_DC_PATCH).astype(int) if ref_keep else np.zeros(n, int)
                d = np.abs(got[k] - ref)
                worst_pix = max(worst_pix, int(d.max())); ndiff += int((d > 0).sum()); total += d.size
        out = {"gate_mean_rel_err": worst_mean, "max_pixel_diff": worst_pix, "pixels_differing": ndiff / max(total, 1), "frequencies_checked": 2}
        assert worst_mean < 1e-5 and worst_pix <= 1
    except AssertionError:
        raise RuntimeError(f"sweep parity failed: {out}")
    return out


# ----------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=DEFAULT_FRAMES, help="frames per GPU per launch")
    ap.add_argument("--arith", default="auto", choices=["auto", "f32", "f32x2"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="headline only")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS), help="only this workload, as a line of its own (development)")
    ap.add_argument("--budget-s", type=float, default=240.0, help="secondary legs are skipped once the run is this old")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    raw = load_capture()
    threads = os.cpu_count() or 1

    # ------------------------------------------------------------ reference arm
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return 0
        kind, frames, run = cpu_reference_rate(raw, threads, target_s=max(2.0, 60.0 / (args.steps + args.warmup)))
        for _ in range(args.warmup):
            run(frames, threads)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run(frames, threads)
        dt = time.perf_counter() - t0
        value = frames * args.steps / dt
        sample = (f"{frames} frames/step of the workload (rf-433.000.raw block replayed), "
                  f"{'reference nrf_fft_new/process/get_buffer via oracle/_ref (FFTW role: in-repo float64 radix-2; libfftw3 absent)' if kind == 'reference' else 'oracle port (float64 radix-2)'}")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "rfdata/rf-433.000.raw (committed fixture), replayed",
            "config": {"workload": "cfg2 nrf_fft path: 1024-pt FFT, all frames of rf-433.000.raw, f32 magnitude + DC patch",
                       "fft_size": N, "hop": HOP, "frames_per_step": frames},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    # ------------------------------------------------------------ our arm
    import torch
    import frequensea_b200 as fb
    ctx = Ctx(args)
    rank, world, local, dev = ctx.rank, ctx.world, ctx.local, ctx.dev
    if fb.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    numa_cpus = fb.binding.host_bind_near_device(local)          # this rank's host thread + pinned buffers next to its GPU

    if args.workload != "cfg2":                                   # development: one workload as its own line
        e = measure_workload(ctx, fb, args.workload, frames_override=None if args.frames == DEFAULT_FRAMES else args.frames,
                             arith={"auto": fb.ARITH_AUTO, "f32": fb.ARITH_F32, "f32x2": fb.ARITH_F32X2}[args.arith])
        if rank == 0:
            w = WORKLOADS[args.workload]
            print(json.dumps({"metric": METRIC, "value": e["value"], "unit": UNIT, "n_gpus": world, "steps": e["steps"], "warmup": args.warmup,
                              "ms_per_step": e["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                              "data": e["data"], "config": {"workload": w["desc"], "fft_size": w["n"], "hop": w["hop"], "kernel": e["kernel"],
                                                           "frames_per_gpu_per_step": e["frames_per_gpu_per_launch"] * e["launches_per_step"]},
                              "parity_gate": e["parity_gate"], "gpu_launches": e["gpu_launches"], "clocks": e["clocks"], "roofline": e["roofline"]}))
        if ctx.clocks:
            ctx.clocks.stop()
        if world > 1:
            import torch.distributed as dist
            dist.barrier(); dist.destroy_process_group()
        return 0

    frames = args.frames - args.frames % FRAMES_PER_BLOCK
    arith = {"auto": fb.ARITH_AUTO, "f32": fb.ARITH_F32, "f32x2": fb.ARITH_F32X2}[args.arith]
    plan = fb.Plan(N, hop=HOP, in_format=fb.IN_S8, out_mode=fb.OUT_MAG, flags=fb.FLAG_DC_PATCH, device=local, arith=arith)

    # host batch: the capture block replayed (rank r starts at a different frame phase)
    reps = frames // FRAMES_PER_BLOCK
    block = np.roll(raw[:BLOCK].reshape(FRAMES_PER_BLOCK, -1), -(rank % FRAMES_PER_BLOCK), axis=0).reshape(-1)
    h_in = torch.empty(frames * N * 2, dtype=torch.uint8, pin_memory=True)
    h_in.view(reps, BLOCK).copy_(torch.from_numpy(block).view(1, BLOCK).expand(reps, BLOCK))
    h_out = torch.empty((frames, N), dtype=torch.float32, pin_memory=True)
    d_in = h_in.to(dev, non_blocking=False)
    d_out = torch.empty((frames, N), dtype=torch.float32, device=dev)
    stream = ctx.stream

    # parity gate before any timing (same build, same data): oracle on a few frames
    plan.exec_device(d_in.data_ptr(), d_out.data_ptr(), frames, stream.cuda_stream)
    torch.cuda.synchronize()
    chk = [0, 1, 77, frames - 1]
    gate, ok = oracle_check(WORKLOADS["cfg2"], d_in, d_out, frames, chk)
    perr = gate["strict_max_err_over_rms"]
    if not ok or not perr < 1e-5:
        raise SystemExit(f"bench.py: parity gate failed ({gate}) -- refusing to time")

    # ---- burst: K single launches, the kernel timed alone (round-1 definition)
    for _ in range(args.warmup):
        plan.exec_device(d_in.data_ptr(), d_out.data_ptr(), frames, stream.cuda_stream)
    b_total, b_steps, (bt0, bt1) = timed_launches(ctx, plan, d_in, d_out, frames, args.steps, 1)
    b_total, b_kernel = ctx.max_over_ranks(b_total, float(np.mean(b_steps)))
    # ---- value: K steps of R launches, >= ~1.2 s under load
    inner = max(1, int(np.ceil(1200.0 / (args.steps * max(b_kernel, 1e-3)))))
    for _ in range(args.warmup):
        plan.exec_device(d_in.data_ptr(), d_out.data_ptr(), frames, stream.cuda_stream)
    launches0 = plan.launches
    total_ms, step_ms, (t0, t1) = timed_launches(ctx, plan, d_in, d_out, frames, args.steps, inner)
    launches = plan.launches - launches0
    clocks = ctx.clocks.region(t0, t1) if ctx.clocks else None
    burst_clocks = ctx.clocks.region(bt0, bt1) if ctx.clocks else None
    total_ms, kernel_ms = ctx.max_over_ranks(total_ms, float(np.mean(step_ms)) / inner)
    value = world * frames * inner * args.steps / (total_ms * 1e-3)

    # ---- e2e: host buffers through the C ABI (H2D + kernel + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, 10))

        def wall(fn, steps):
            for _ in range(2):
                fn()
            ctx.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                fn()
            torch.cuda.synchronize()
            return ctx.max_over_ranks(time.perf_counter() - t0)[0]

        dt = wall(lambda: plan.exec_host_ptr(h_in.data_ptr(), h_out.data_ptr(), frames), e2e_steps)
        e2e_ok = bool(torch.equal(h_out[chk], d_out[chk].cpu()))
        dt_copy = wall(lambda: plan.copy_only_host_ptr(h_in.data_ptr(), h_out.data_ptr(), frames), e2e_steps)
        ib, ob = plan.input_bytes(frames), plan.output_bytes(frames)
        e2e = {"value": world * frames * e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": world * ib, "d2h_bytes_per_step": world * ob,
               "steps": e2e_steps, "host_memory": "pinned", "matches_device_path": e2e_ok,
               "gbs_h2d_plus_d2h": world * (ib + ob) * e2e_steps / dt / 1e9,
               "copy_only": {"value": world * frames * e2e_steps / dt_copy, "unit": UNIT, "gbs_h2d_plus_d2h": world * (ib + ob) * e2e_steps / dt_copy / 1e9,
                             "note": "fsea_copy_only_host: the same chunks over the same three streams, kernel left out -- the PCIe / host-memory ceiling of this job on this box"},
               "frac_of_copy_only": dt_copy / dt,
               "numa": f"rank thread + pinned buffers bound to the GPU's node ({numa_cpus} CPUs)" if numa_cpus > 0 else "topology unknown: not bound"}
        # the sweep's product form: u8 pixels out (3N bytes per frame instead of 6N)
        try:
            pplan = fb.Plan(N, hop=HOP, in_format=fb.IN_S8, out_mode=fb.OUT_PIX, pix_scale=10.0, device=local)
            hp = torch.empty((frames, N), dtype=torch.uint8, pin_memory=True)
            dtp = wall(lambda: pplan.exec_host_ptr(h_in.data_ptr(), hp.data_ptr(), frames), e2e_steps)
            e2e["pix_u8"] = {"value": world * frames * e2e_steps / dtp, "unit": UNIT, "d2h_bytes_per_step": world * frames * N,
                             "gbs_h2d_plus_d2h": world * (ib + frames * N) * e2e_steps / dtp / 1e9}
            del pplan, hp
        except Exception as e:
            e2e["pix_u8"] = {"error": str(e)[:200]}

    del h_out, d_out, d_in, h_in
    torch.cuda.empty_cache()

    # ---- secondary workloads (every rank measures; rank 0 reports)
    secondary, sustained, sweep_res = {}, None, None
    if not args.no_secondary:
        names = SECONDARY_ORDER if world == 1 else ["cfg5", "cfg3", "cfg4"]
        for name in names:
            if ctx.max_over_ranks(ctx.elapsed())[0] > args.budget_s:
                secondary[name] = {"skipped": "time budget"}
                continue
            try:
                secondary[name] = measure_workload(ctx, fb, name)
            except Exception as e:
                secondary[name] = {"error": str(e)[:300]}
        if world == 1 and ctx.elapsed() <= args.budget_s:
            # cfg4's rows in the reference's own arithmetic (FSEA_ARITH_F64): the pixels must EQUAL the oracle's
            try:
                ex = measure_workload(ctx, fb, "cfg4", frames_override=512, min_region_s=0.2, arith=fb.ARITH_F64)
                ex["note"] = ("FSEA_ARITH_F64: samples widened in double on the device, fp64 transform; the mode whose pixels are the "
                              "reference's byte for byte (roofline fraction is not the point of this entry)")
                if ex["parity_gate"]["max_pixel_diff"] != 0:
                    raise RuntimeError(f"fp64 pixels differ from the oracle: {ex['parity_gate']}")
                secondary["cfg4_pixels_f64"] = ex
            except Exception as e:
                secondary["cfg4_pixels_f64"] = {"error": str(e)[:300]}
        if world == 1 and ctx.elapsed() <= args.budget_s:
            try:
                secondary["file_ingest"] = file_ingest(ctx, fb)
            except Exception as e:
                secondary["file_ingest"] = {"error": str(e)[:300]}
        if ctx.max_over_ranks(ctx.elapsed())[0] <= args.budget_s:
            try:
                sustained = sustained_cfg5(ctx, fb, seconds=5.0)
            except Exception as e:
                sustained = {"error": str(e)[:300]}
        if ctx.max_over_ranks(ctx.elapsed())[0] <= args.budget_s + 30:
            try:
                sweep_res = sweep_job(ctx, fb)
            except Exception as e:
                sweep_res = {"error": str(e)[:300]}
        secondary["sustained"] = sustained
        secondary["sweep"] = sweep_res

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the (single) kernel
    peak, peak_src = peak_hbm()
    alg_bytes = frames * (2 * HOP + 4 * N)
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    burst_achieved = alg_bytes / (b_kernel * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if tj.get("frames") == frames and tj.get("fft_size") == N:
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": plan.kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": "profiles/ncu_traffic.json (one ncu --set full capture of this kernel and batch, not this run)" if traffic else None,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms, "peak_source": peak_src,
                "burst_frac": burst_achieved / peak, "burst_kernel_ms": b_kernel,
                "note": "frac: average launch inside the >= 1.2 s timed region; burst_frac: K single launches (7 ms), the kernel timed alone"}

    # ---- cpu baseline (rank 0, N=1 only): bounded sample of the same frames
    cpu = None
    if world == 1 and not args.no_cpu:
        kind, cframes, run = cpu_reference_rate(raw, threads, target_s=12.0)
        dt = run(cframes, threads)
        cpu = {"value": cframes / dt, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"{cframes} frames of the same workload in {dt:.1f} s, {threads} threads; "
                         + ("reference nrf_fft_* via oracle/_ref (FFTW role: in-repo float64 radix-2, libfftw3 absent)" if kind == "reference" else "oracle port")}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "rfdata/rf-433.000.raw capture block (committed fixture) replayed to batch size",
        "config": {"workload": "cfg2 nrf_fft path: 1024-pt FFT, all frames of rf-433.000.raw, int8 IQ -> f32 magnitude + DC patch",
                   "fft_size": N, "hop": HOP, "frames_per_gpu_per_step": frames * inner, "launches_per_step": inner, "frames_per_gpu_per_launch": frames,
                   "parallelism": f"frames sharded over {world} GPU(s), no collective",
                   "l2": "inputs larger than L2: 512 MiB in + 1 GiB out per launch per GPU" if frames >= (1 << 17) else "WARNING: batch smaller than L2",
                   "kernel": plan.kernel_name, "arith": args.arith},
        "iq_gsps": value * HOP / 1e9, "parity_max_rel_err": perr, "parity_gate": gate,
        "burst": {"value": world * frames * args.steps / (b_total * 1e-3), "unit": UNIT, "ms_per_launch": b_kernel, "launches": args.steps, "clocks": burst_clocks},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "secondary": secondary if not args.no_secondary else None,
        "bench_seconds": ctx.elapsed(),
    }
    print(json.dumps(line))
    if ctx.clocks:
        ctx.clocks.stop()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
