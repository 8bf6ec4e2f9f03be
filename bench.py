#!/usr/bin/env python3
"""bench.py -- headline benchmark of the IQ -> spectrum hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): the nrf_fft path --
1024-point FFT of every frame of rfdata/rf-433.000.raw (two's-complement int8
IQ, 128 frames per 262144-byte block), float32 magnitude with the reference's
DC patch -- with the capture block replayed R times so that one step's input
(512 MiB) and output (1 GiB) are far larger than L2 (126 MB).  One "step" =
one pass of the hot path over that batch.  Metric: frames/s (whole job).

  value     device-resident batch (fsea_exec_device), CUDA-event timed, max over ranks
  e2e       same batch through the C ABI with HOST buffers (fsea_exec_host):
            pinned host -> H2D -> kernel -> D2H -> pinned host, wall clock
  roofline  algorithmic bytes (2*hop + 4*N per frame) / kernel time
            against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own nrf_fft_* (oracle/_ref) or the oracle port,
            timed on this box's host cores on a bounded sample of the same frames

Multi-GPU (torchrun, one rank per GPU): frames are independent, so each rank
transforms its own shard -- no data-path collective; weak scaling.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N = 1024
HOP = 1024
FRAMES_PER_BLOCK = 128
BLOCK = 262144
DEFAULT_FRAMES = 1 << 18          # per GPU per step: 512 MiB in, 1 GiB out
METRIC = "fft_frames_per_s"
UNIT = "frames/s"

# Secondary workloads; the headline line is always cfg2.  Synthetic IQ is
# drawn on the device (uniform bytes): throughput does not depend on the sample values and
# the parity gate checks the very frames that are timed.
def _size_workload(n):
    frames = max(256, (1 << 28) // n)            # 512 MiB of raw IQ per step
    return dict(n=n, hop=n, fmt="s8", window=False, out="mag", patch=True, frames=frames,
                desc=f"size ladder: {n}-pt FFT, int8 IQ -> f32 magnitude + DC patch")


WORKLOADS = {
    "cfg2": dict(n=1024, hop=1024, fmt="s8", window=False, out="mag", patch=True, frames=1 << 18,
                 desc="cfg2 nrf_fft path: 1024-pt FFT, all frames of rf-433.000.raw, int8 IQ -> f32 magnitude + DC patch"),
    "cfg3": dict(n=8192, hop=4096, fmt="u8", window=True, out="logmag", patch=False, frames=65535,
                 desc="cfg3: 8192-pt FFT, Hann, 50% overlap, synthetic uint8 IQ -> f32 log-magnitude"),
    "cfg4": dict(n=16384, hop=16384, fmt="s8", window=False, out="pix", patch=True, frames=1 << 14, scale=5.0,
                 desc="cfg4 sweep rows: 16384-pt FFT, int8 IQ -> u8 pixels (S=5, DC patch)"),
    "cfg5": dict(n=65536, hop=65536, fmt="s8", window=False, out="logmag", patch=False, frames=4096,
                 desc="cfg5: 65536-pt FFT, synthetic int8 HackRF IQ -> f32 log-magnitude"),
}
for _n in (128, 256, 512, 2048, 4096, 8192, 16384, 32768, 65536):
    WORKLOADS[f"n{_n}"] = _size_workload(_n)


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
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(",") for r in open(self.tmp.name) if r.strip()]
        os.unlink(self.tmp.name)
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                r = [x.strip() for x in r]
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            hot = [s for s, p in zip(sm, power) if p >= 0.6 * max(power)] or sm
            out.update(sm_mhz=float(np.median(hot)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(power)))
        return out


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


def cpu_reference_rate(raw, threads, target_s=12.0, impl_only=False):
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


def secondary(args):
    """Device-resident throughput + roofline of a secondary workload; under torchrun every
    rank runs the same per-GPU batch on its own device (independent frames / sweep
    frequencies: no collective, weak scaling), time = max over ranks."""
    import torch
    import frequensea_b200 as fb
    from oracle import orc
    rank, world, local = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    w = WORKLOADS[args.workload]
    n, hop, frames = w["n"], w["hop"], w["frames"]
    if args.frames != DEFAULT_FRAMES:
        frames = args.frames
    fmt = fb.IN_S8 if w["fmt"] == "s8" else fb.IN_U8
    out_mode = {"mag": fb.OUT_MAG, "logmag": fb.OUT_LOGMAG, "pix": fb.OUT_PIX}[w["out"]]
    flags = fb.FLAG_DC_PATCH if w["patch"] else 0
    arith = {"auto": fb.ARITH_AUTO, "f32": fb.ARITH_F32, "f32x2": fb.ARITH_F32X2}[args.arith]
    plan = fb.Plan(n, hop=hop, in_format=fmt, window=fb.WIN_HANN if w["window"] else fb.WIN_NONE,
                   out_mode=out_mode, flags=flags, pix_scale=w.get("scale", 10.0), arith=arith, device=local)
    dev = torch.device("cuda", local)
    g = torch.Generator(device=dev).manual_seed(20260925 + rank)
    d_in = torch.randint(0, 256, (plan.input_bytes(frames),), dtype=torch.uint8, device=dev, generator=g)
    d_out = torch.empty((frames, n), dtype=torch.uint8 if out_mode == fb.OUT_PIX else torch.float32, device=dev)
    stream = torch.cuda.current_stream()
    plan.exec_device(d_in.data_ptr(), d_out.data_ptr(), frames, stream.cuda_stream)
    torch.cuda.synchronize()
    # parity gate on the timed data
    chk = [0, 1, frames // 2, frames - 1]
    worst = 0.0
    for f in chk:
        raw = d_in[f * hop * 2:f * hop * 2 + 2 * n].cpu().numpy()
        ref = orc.engine_frame(raw, n, orc.IN_S8 if w["fmt"] == "s8" else orc.IN_U8,
                               orc.WIN_HANN if w["window"] else orc.WIN_NONE,
                               {"mag": orc.OUT_MAG, "logmag": orc.OUT_LOGMAG, "pix": orc.OUT_PIX}[w["out"]],
                               w.get("scale", 10.0), orc.FLAG_DC_PATCH if w["patch"] else 0).astype(np.float64)
        got = d_out[f].cpu().numpy().astype(np.float64)
        if w["out"] == "mag":
            worst = max(worst, float(np.max(np.abs(got - ref)) / np.sqrt(np.mean(ref ** 2))))
            ok = worst < 1e-5
        elif w["out"] == "logmag":
            lin = 10 ** (ref / 20)
            m = lin >= 1e-3 * np.sqrt(np.mean(lin ** 2))
            worst = max(worst, float(np.max(np.abs(got - ref)[m])))
            ok = worst < 2e-3
        else:
            worst = max(worst, float(np.max(np.abs(got - ref))))
            ok = worst <= 1
        if not ok:
            raise SystemExit(f"bench.py: parity gate failed on {args.workload} ({worst})")
    for _ in range(args.warmup):
        plan.exec_device(d_in.data_ptr(), d_out.data_ptr(), frames, stream.cuda_stream)
    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = plan.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        plan.exec_device(d_in.data_ptr(), d_out.data_ptr(), frames, stream.cuda_stream)
    e1.record(stream)
    barrier()
    tms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms[0])
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()
        return 0
    peak, peak_src = peak_hbm()
    out_b = 1 if w["out"] == "pix" else 4
    alg = frames * (2 * hop + out_b * n)
    achieved = alg / (ms * 1e-3) / 1e9
    # end to end from a capture FILE (raw 8-bit IQ as c/rfcap.c / c/reader.c write it): pread -> pinned
    # slots -> H2D -> kernels -> D2H, through fsea_exec_file; rank 0, N = 1 only
    file_ingest = None
    if world == 1 and not args.no_e2e:
        try:
            fframes = min(frames, max(2, (1 << 29) // (2 * hop)))            # <= 512 MiB of capture
            shm = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
            path = os.path.join(shm, f"fsea_bench_{os.getpid()}.raw")
            d_in[:plan.input_bytes(fframes)].cpu().numpy().tofile(path)
            t_out = torch.empty((fframes, n), dtype=torch.uint8 if out_mode == fb.OUT_PIX else torch.float32).pin_memory()
            h_out = t_out.numpy()
            plan.exec_file(path, fframes, 0, out=h_out)                       # warm-up (page cache, staging slots)
            assert np.array_equal(h_out[fframes - 1], d_out[fframes - 1].cpu().numpy()), "file path differs from device path"
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                plan.exec_file(path, fframes, 0, out=h_out)
            dt = (time.perf_counter() - t0) / reps
            os.unlink(path)
            file_ingest = {"value": fframes / dt, "unit": UNIT, "msps": fframes * hop / dt / 1e6, "file_bytes": plan.input_bytes(fframes),
                           "d2h_bytes": int(h_out.nbytes), "source": f"capture file on {shm} (page cache), pinned result buffer",
                           "matches_device_path": True}
        except Exception as e:                                                # never fail the bench line over the extra leg
            file_ingest = {"error": str(e)[:200]}
    print(json.dumps({
        "metric": METRIC, "value": world * frames / (ms * 1e-3), "file_ingest": file_ingest, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic uniform 8-bit IQ drawn on the device",
        "config": {"workload": w["desc"], "fft_size": n, "hop": hop, "frames_per_gpu_per_step": frames, "kernel": plan.kernel_name,
                   "parallelism": f"frames / sweep frequencies sharded over {world} GPU(s), no collective",
                   "l2": f"inputs larger than L2: {plan.input_bytes(frames) >> 20} MiB in + {plan.output_bytes(frames) >> 20} MiB out per step"},
        "iq_gsps": world * frames * hop / (ms * 1e-3) / 1e9, "parity_gate": worst, "gpu_launches": plan.launches - l0, "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": plan.kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "algorithmic_bytes_per_step": alg, "peak_source": peak_src,
                     "note": "per GPU, per step (all launches of the step)"},
    }))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=DEFAULT_FRAMES, help="frames per GPU per step")
    ap.add_argument("--arith", default="auto", choices=["auto", "f32", "f32x2"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    if args.workload != "cfg2":
        return secondary(args)
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
    rank, world, local = dist_setup(args.gpus)
    if fb.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
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
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # parity gate before any timing (same build, same data): oracle on a few frames
    from oracle import orc
    plan.exec_device(d_in.data_ptr(), d_out.data_ptr(), frames, stream.cuda_stream)
    torch.cuda.synchronize()
    chk = [0, 1, 77, frames - 1]
    got = d_out[chk].cpu().numpy().astype(np.float64)
    hin = h_in.numpy()
    ref = np.stack([orc.engine_frame(hin[f * 2 * N:], N, orc.IN_S8, orc.WIN_NONE, orc.OUT_MAG, flags=orc.FLAG_DC_PATCH) for f in chk])
    perr = float(np.max(np.abs(got - ref) / np.sqrt(np.mean(ref ** 2, axis=1, keepdims=True))))
    if not perr < 1e-5:
        raise SystemExit(f"bench.py: parity gate failed (max rel err {perr:.3e}) -- refusing to time")

    # ---- value: device-resident, CUDA events on the launching stream
    for _ in range(args.warmup):
        plan.exec_device(d_in.data_ptr(), d_out.data_ptr(), frames, stream.cuda_stream)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = plan.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for a, b in ev:
        a.record(stream)
        plan.exec_device(d_in.data_ptr(), d_out.data_ptr(), frames, stream.cuda_stream)
        b.record(stream)
    e1.record(stream)
    barrier()
    total_ms = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    launches = plan.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms_max = float(t[0]), float(t[1])
    value = world * frames * args.steps / (total_ms * 1e-3)

    # ---- e2e: host buffers through the C ABI (H2D + kernel + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            plan.exec_host_ptr(h_in.data_ptr(), h_out.data_ptr(), frames)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            plan.exec_host_ptr(h_in.data_ptr(), h_out.data_ptr(), frames)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ok = bool(torch.equal(h_out[chk], d_out[chk].cpu()))
        e2e = {"value": world * frames * e2e_steps / float(t[0]), "unit": UNIT,
               "h2d_bytes_per_step": world * plan.input_bytes(frames), "d2h_bytes_per_step": world * plan.output_bytes(frames),
               "steps": e2e_steps, "host_memory": "pinned", "matches_device_path": e2e_ok}

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the (single) kernel
    peak, peak_src = peak_hbm()
    alg_bytes = frames * (2 * HOP + 4 * N)
    achieved = alg_bytes / (kernel_ms_max * 1e-3) / 1e9
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
                "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_ms": kernel_ms_max, "peak_source": peak_src}

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
                   "fft_size": N, "hop": HOP, "frames_per_gpu_per_step": frames, "parallelism": f"frames sharded over {world} GPU(s), no collective",
                   "l2": "inputs larger than L2: 512 MiB in + 1 GiB out per step per GPU" if frames >= (1 << 17) else "WARNING: batch smaller than L2",
                   "kernel": plan.kernel_name, "arith": args.arith},
        "iq_gsps": value * HOP / 1e9, "parity_max_rel_err": perr,
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
