#!/usr/bin/env python
"""BASELINE cfg4 through the one-process multi-GPU driver (libfsea_mgpu.so: ncclCommInitAll, G devices).

  python tools/sweep_mgpu.py --gpus 8 [--freqs 1000] [--rows 512] [--fft-size 16384] [--mode gather|p2p|local]
                             [--check]   (--check: the G-GPU job must equal the 1-GPU job bit for bit)

No torch in this process: the driver links the system NCCL, torch ships another.
Prints one JSON line: rows/s of the whole job, per-phase device times, SURVEY 8(e) numbers.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Cfg(C.Structure):
    _fields_ = [("n_gpus", C.c_int), ("fft_size", C.c_int), ("rows", C.c_int), ("n_freqs", C.c_int), ("gate_rows", C.c_int),
                ("gate_threshold", C.c_double), ("pix_scale", C.c_float), ("width_step", C.c_int), ("quiet_every", C.c_int),
                ("mode", C.c_int), ("reserved", C.c_int * 6)]


class Timing(C.Structure):
    _fields_ = [("stage_s", C.c_double), ("gate_s", C.c_double), ("render_s", C.c_double), ("exchange_s", C.c_double),
                ("stitch_s", C.c_double), ("total_s", C.c_double), ("rows_rendered", C.c_long), ("kept", C.c_long),
                ("launches", C.c_long), ("reserved", C.c_double * 4)]


MODES = {"gather": 0, "p2p": 1, "local": 2}


def load():
    C.CDLL(os.path.join(ROOT, "frequensea_b200", "libfsea.so"), mode=C.RTLD_GLOBAL)
    L = C.CDLL(os.path.join(ROOT, "frequensea_b200", "libfsea_mgpu.so"))
    L.fsea_mgpu_sweep.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Timing)]
    L.fsea_mgpu_sweep.restype = C.c_int
    L.fsea_mgpu_gather_bench.argtypes = [C.c_int, C.c_long, C.c_int, C.POINTER(C.c_double)]
    L.fsea_mgpu_gather_bench.restype = C.c_int
    L.fsea_mgpu_last_error.restype = C.c_char_p
    return L


def run(L, gpus, n, rows, freqs, mode, step, want_picture, gate_rows=100, quiet_every=10):
    cfg = Cfg(n_gpus=gpus, fft_size=n, rows=rows, n_freqs=freqs, gate_rows=gate_rows, gate_threshold=1.1, pix_scale=5.0,
              width_step=step, quiet_every=quiet_every, mode=MODES[mode])
    t = Timing()
    means = np.zeros(freqs, dtype=np.float64)
    keep = np.zeros(freqs, dtype=np.uint8)
    pic = np.zeros((rows, n + (freqs - 1) * step), dtype=np.uint8) if want_picture else None
    rc = L.fsea_mgpu_sweep(C.byref(cfg), pic.ctypes.data_as(C.c_void_p) if pic is not None else None,
                           means.ctypes.data_as(C.c_void_p), keep.ctypes.data_as(C.c_void_p), C.byref(t))
    if rc != 0:
        raise RuntimeError(f"fsea_mgpu_sweep failed ({rc}): {L.fsea_mgpu_last_error().decode()}")
    return means, keep, pic, t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--freqs", type=int, default=1000)
    ap.add_argument("--rows", type=int, default=512)
    ap.add_argument("--fft-size", dest="n", type=int, default=16384)
    ap.add_argument("--mode", default="gather", choices=sorted(MODES))
    ap.add_argument("--step", type=int, default=0, help="WIDTH_STEP (0: N, the c/fft-stitch-broad.c convention)")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--out", default="", help="write the stitched picture as a binary PGM (P5): what c/fft-stitch-broad.c hands to its PNG writer")
    ap.add_argument("--f32-gather-mib", type=int, default=1024, help="SURVEY 8(e) (iii): MiB each GPU sends to the root (0: skip)")
    a = ap.parse_args()
    L = load()
    step = a.step or a.n
    picture = a.check or bool(a.out)
    means, keep, pic, t = run(L, a.gpus, a.n, a.rows, a.freqs, a.mode, step, picture)
    out = {"workload": "cfg4 sweep job (one process, ncclCommInitAll)", "n_gpus": a.gpus, "freqs": a.freqs, "rows": a.rows,
           "fft_size": a.n, "mode": a.mode, "width_step": step, "kept": int(t.kept), "rows_rendered": int(t.rows_rendered),
           "seconds": {"stage_untimed": t.stage_s, "gate": t.gate_s, "render": t.render_s, "exchange": t.exchange_s,
                       "stitch": t.stitch_s, "total": t.total_s},
           "rows_per_s": t.rows_rendered / t.total_s if t.total_s > 0 else None,
           "render_rows_per_s_per_gpu": (t.rows_rendered / a.gpus) / t.render_s if t.render_s > 0 else None,
           "gpu_launches": int(t.launches),
           "image_bytes_exchanged": int(t.kept) * a.rows * a.n * (a.gpus - 1) // max(a.gpus, 1) if a.mode != "local" else 0}
    if t.exchange_s > 0:
        out["exchange_gbs_into_root"] = out["image_bytes_exchanged"] / t.exchange_s / 1e9
    if a.check and a.gpus > 1:
        m1, k1, p1, _ = run(L, 1, a.n, a.rows, a.freqs, "gather", step, True)
        out["equals_one_gpu_job"] = bool(np.array_equal(means, m1) and np.array_equal(keep, k1) and (pic is None or np.array_equal(pic, p1)))
        out["picture_sha256"] = hashlib.sha256(pic.tobytes()).hexdigest()[:16] if pic is not None else None
    if a.gpus > 1 and a.f32_gather_mib > 0:
        sec = C.c_double()
        rc = L.fsea_mgpu_gather_bench(a.gpus, a.f32_gather_mib << 20, 3, C.byref(sec))
        if rc == 0:
            total = (a.gpus - 1) * (a.f32_gather_mib << 20)
            out["f32_gather_to_root"] = {"bytes": total, "seconds": sec.value, "gbs_into_root": total / sec.value / 1e9,
                                         "note": "SURVEY 8(e) (iii): float32 spectra gathered on one root are bound by its NVLink ingest"}
    if a.out and pic is not None:
        with open(a.out, "wb") as fh:                      # rows x width u8, row-major: a PGM is the header plus exactly these bytes
            fh.write(b"P5\n%d %d\n255\n" % (pic.shape[1], pic.shape[0]))
            fh.write(np.ascontiguousarray(pic).tobytes())
        out["picture_file"] = {"path": a.out, "width": int(pic.shape[1]), "height": int(pic.shape[0])}
    print(json.dumps(out))
    if a.check and a.gpus > 1:
        assert out["equals_one_gpu_job"]


if __name__ == "__main__":
    main()
