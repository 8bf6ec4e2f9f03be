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
