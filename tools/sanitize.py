#!/usr/bin/env python
"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tools/sanitize.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import frequensea_b200 as fb  # noqa: E402
from conftest import synth_s8, synth_u8  # noqa: E402

CASES = [
    (64, 9, {}), (128, 70, {}), (1024, 5, {}), (1024, 5, {"window": fb.WIN_HANN, "hop": 512}),
    (2048, 3, {}), (2048, 3, {"arith": fb.ARITH_F32X2}), (4096, 3, {}), (4096, 2, {"arith": fb.ARITH_F32}),
    (8192, 3, {}), (8192, 3, {"window": fb.WIN_HANN, "hop": 4096, "out_mode": fb.OUT_LOGMAG}),
    (8192, 2, {"arith": fb.ARITH_F32X2}),
    (16384, 5, {"out_mode": fb.OUT_PIX, "flags": fb.FLAG_DC_PATCH, "pix_scale": 5.0}),
    (32768, 3, {"out_mode": fb.OUT_LOGMAG}), (65536, 3, {}),
]
for n, frames, kw in CASES:
    hop = kw.get("hop", n)
    data = synth_s8((frames - 1) * hop + n, seed=n) if "window" not in kw else synth_u8((frames - 1) * hop + n, seed=n)
    plan = fb.Plan(n, in_format=fb.IN_S8 if "window" not in kw else fb.IN_U8, **kw)
    out = plan.exec_host(data, frames)
    print(n, plan.kernel_name, out.shape, float(np.asarray(out, dtype=np.float64).sum()))

# round 2: sweep engine (side streams, grouped output rows, device-drawn source), fused front chain, device history ring,
# float frames above 8192 points, and the cluster form of the warp-specialised kernel at 65536 (TMA, mbarriers, DSMEM)
from frequensea_b200.binding import Chain, Sweep  # noqa: E402
eng = Sweep(4096, 12, 5, gate_rows=4, own_images=True)
for f in range(5):
    eng.submit_synth(f, f, 3)
means, keep = eng.gate(5)
eng.render(5)
print("sweep", keep.tolist(), int(eng.read_images(0, 5).sum()))
nrf = fb.nrf
blk = synth_u8(131072, seed=3)
coef = nrf.fir_low_pass_coefficients(5000000, 200000, 51)[:51]
ch = Chain(1024, shifter=(-10000, 5000000), fir_coefficients=coef)
print("chain", float(ch.process(blk).sum()), float(ch.process(blk).sum()))
fft = nrf.fft_new(1024, 8)
for _ in range(10):
    nrf.fft_process(fft, nrf.nut_buffer_new_u8(131072, 2, blk))
nrf.fft_shift(fft, 16.0)
print("history", float(nrf.fft_get_buffer(fft).numpy().sum()))
nrf.fft_free(fft)
x = np.random.default_rng(1).normal(0, 1, 2 * 16384)
print("bigf", float(fb.Plan(16384, in_format=fb.IN_F64C).exec_host(x, 1).sum()))
# float64 arithmetic on 8-bit input (widen kernel + fp64 register kernel / radix-16 global passes) and the sweep gate at
# N = 16384 (per-launch hop through the tensor map, SUMS form of the warp-specialised kernel), pixel pass in fp64
for n in (1024, 32768):
    pl = fb.Plan(n, hop=n // 2, in_format=fb.IN_U8, window=fb.WIN_HANN, out_mode=fb.OUT_PIX, pix_scale=5.0, flags=fb.FLAG_DC_PATCH, arith=fb.ARITH_F64)
    print(pl.kernel_name, int(pl.exec_host(synth_u8(2 * n, seed=n + 1), 3).sum()))
eng = Sweep(16384, 6, 4, gate_rows=3, own_images=True, pixel_arith=fb.ARITH_F64)
for f in range(4):
    eng.submit_synth(f, f, 3)
means, keep = eng.gate(4)
eng.render(4)
print("sweep 16384 (sums gate, f64 pixels)", keep.tolist(), [round(float(m), 4) for m in means], int(eng.read_images(0, 4).sum()))
os.environ["FSEA_WS"] = "1"
plan = fb.Plan(65536, in_format=fb.IN_S8, out_mode=fb.OUT_LOGMAG)
print(plan.kernel_name, float(plan.exec_host(synth_s8(3 * 65536, seed=9), 3).sum()))
