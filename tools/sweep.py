#!/usr/bin/env python
"""Kernel-only A/B sweep in ONE process (development tool, not the bench).

  python tools/sweep.py n2048 n2048:f32 cfg3 n8192::build_alt/libfsea_x.so ...

Each item is workload[@frames][:arith[:lib[:ENV=VAL,...]]] with the workloads of bench.py.  Inputs are
device-resident and larger than L2; 3 warm-ups, 5 timed launches, CUDA events;
a light parity spot-check (oracle, 2 frames) guards against timing a broken build.
Prints one line per item: frames/s and fraction of the algorithmic HBM roofline.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from frequensea_b200 import binding as B  # noqa: E402
from oracle import orc  # noqa: E402

_libs = {}


def load(path):
    if path not in _libs:
        L = C.CDLL(path)
        vp, i, l = C.c_void_p, C.c_int, C.c_long
        P = C.POINTER(B.FseaConfig)
        L.fsea_config_default.argtypes = [P, i]; L.fsea_config_default.restype = None
        L.fsea_plan_create.argtypes = [P, C.POINTER(vp)]; L.fsea_plan_create.restype = i
        L.fsea_plan_destroy.argtypes = [vp]
        L.fsea_input_bytes.argtypes = [vp, l]; L.fsea_input_bytes.restype = C.c_size_t
        L.fsea_exec_device.argtypes = [vp, vp, vp, l, vp]; L.fsea_exec_device.restype = i
        L.fsea_plan_kernel_name.argtypes = [vp]; L.fsea_plan_kernel_name.restype = C.c_char_p
        L.fsea_plan_launch_count.argtypes = [vp]; L.fsea_plan_launch_count.restype = l
        L.fsea_last_cuda_error.restype = C.c_char_p
        _libs[path] = L
    return _libs[path]


def run(item, steps=int(os.environ.get("SWEEP_STEPS", "5")), warmup=int(os.environ.get("SWEEP_WARMUP", "3"))):
    parts = (item.split(":") + ["", "", ""])[:4]
    wl, arith, libp = parts[0], parts[1] or "auto", parts[2] or B.lib_path()
    envs = [kv.split("=") for kv in parts[3].split(",") if kv]          # ENV=VAL[,ENV=VAL] for this item only
    for k, v in envs:
        os.environ[k] = v
    frames_override = None
    if "@" in wl:
        wl, fo = wl.split("@"); frames_override = int(fo)
    w = bench.WORKLOADS[wl]
    n, hop, frames = w["n"], w["hop"], frames_override or w["frames"]
    L = load(os.path.join(ROOT, libp) if not os.path.isabs(libp) else libp)
    cfg = B.FseaConfig()
    L.fsea_config_default(C.byref(cfg), n)
    cfg.hop = hop
    cfg.in_format = B.IN_S8 if w["fmt"] == "s8" else B.IN_U8
    cfg.window = B.WIN_HANN if w["window"] else B.WIN_NONE
    cfg.out_mode = {"mag": B.OUT_MAG, "logmag": B.OUT_LOGMAG, "pix": B.OUT_PIX}[w["out"]]
    cfg.flags = B.FLAG_DC_PATCH if w["patch"] else 0
    cfg.pix_scale = w.get("scale", 10.0)
    cfg.arith = {"auto": 0, "f32": 1, "f32x2": 2, "f32x2h": 3, "r64": 4, "f64": 5}[arith]
    h = C.c_void_p()
    rc = L.fsea_plan_create(C.byref(cfg), C.byref(h))
    if rc:
        print(f"{item:40s} plan_create failed {rc} {L.fsea_last_cuda_error().decode()}")
        return
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(7)
    nbytes = L.fsea_input_bytes(h, frames)
    d_in = torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device=dev, generator=g)
    pix = cfg.out_mode == B.OUT_PIX
    d_out = torch.zeros((frames, n), dtype=torch.uint8 if pix else (torch.float64 if arith == "f64" else torch.float32), device=dev)
    s = torch.cuda.current_stream().cuda_stream

    def go():
        rc = L.fsea_exec_device(h, d_in.data_ptr(), d_out.data_ptr(), frames, s)
        if rc:
            raise RuntimeError(f"exec {rc} {L.fsea_last_cuda_error().decode()}")
    go(); torch.cuda.synchronize()
    worst = 0.0
    for f in (1, frames - 1):
        raw = d_in[f * hop * 2:f * hop * 2 + 2 * n].cpu().numpy()
        ref = orc.engine_frame(raw, n, orc.IN_S8 if w["fmt"] == "s8" else orc.IN_U8,
                               orc.WIN_HANN if w["window"] else orc.WIN_NONE,
                               {"mag": orc.OUT_MAG, "logmag": orc.OUT_LOGMAG, "pix": orc.OUT_PIX}[w["out"]],
                               w.get("scale", 10.0), orc.FLAG_DC_PATCH if w["patch"] else 0).astype(np.float64)
        got = d_out[f].cpu().numpy().astype(np.float64)
        if w["out"] == "mag":
            worst = max(worst, float(np.max(np.abs(got - ref)) / np.sqrt(np.mean(ref ** 2))))
        elif w["out"] == "logmag":
            lin = 10 ** (ref / 20); m = lin >= 1e-3 * np.sqrt(np.mean(lin ** 2))
            worst = max(worst, float(np.max(np.abs(got - ref)[m])))
        else:
            worst = max(worst, float(np.max(np.abs(got - ref))))
    for _ in range(warmup):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    fps = frames / (ms * 1e-3)
    alg = frames * (hop * 2 + n * (1 if pix else 4))
    peak, _ = bench.peak_hbm()
    print(f"{item:40s} {L.fsea_plan_kernel_name(h).decode():34s} {fps / 1e6:10.2f} Mframes/s  frac {alg / (ms * 1e-3) / 1e9 / peak:.3f}  err {worst:.2e}", flush=True)
    L.fsea_plan_destroy(h)
    for k, _ in envs:
        os.environ.pop(k, None)
    del d_in, d_out
    torch.cuda.empty_cache()


if __name__ == "__main__":
    for it in sys.argv[1:]:
        try:
            run(it)
        except Exception as e:  # keep sweeping
            print(f"{it:40s} ERROR {e}", flush=True)
