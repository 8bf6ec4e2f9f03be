#!/usr/bin/env python
"""Per-draw cost of the reference-facing operators, as lua/fft.lua:42-50 calls them once per frame:

    nrf_fft_process(fft, samples_buffer); fft_buffer = nrf_fft_get_buffer(fft)

libfsea.so (one frame through fsea_exec_host + the host history ring) next to the compiled
reference (oracle/_ref: 2 x 131072 powf calls, FFT, 8 MiB scroll, 8 MiB copy, src/nrf.c:598-635).
Prints one JSON line.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import frequensea_b200 as fb  # noqa: E402
from frequensea_b200 import nrf  # noqa: E402
from oracle import orc  # noqa: E402


def main():
    z = np.load(os.path.join(ROOT, "tests", "golden", "iq_fixtures.npz"))
    u8 = orc.device_unpack(z["rf433"])
    out = {"metric": "operator_us_per_draw", "unit": "us", "shapes": {}}
    for n, hist in ((1024, 1024), (128, 512), (128, 128)):
        fft = nrf.fft_new(n, hist)
        buf = nrf.nut_buffer_new_u8(131072, 2, u8)
        for _ in range(20):
            nrf.fft_process(fft, buf); nrf.fft_get_buffer(fft).free()
        t0 = time.perf_counter(); reps = 300
        for _ in range(reps):
            nrf.fft_process(fft, buf)
        t1 = time.perf_counter()
        for _ in range(reps):
            nrf.fft_process(fft, buf); nrf.fft_get_buffer(fft).free()
        t2 = time.perf_counter()
        ours_p, ours_pg = (t1 - t0) / reps * 1e6, (t2 - t1) / reps * 1e6
        last = nrf.fft_get_buffer(fft).numpy()[:n]
        nrf.fft_free(fft)
        R = orc.ref()
        rfft = R.nrf_fft_new(n, hist)
        nb = R.nut_buffer_new_u8(131072, 2, u8.ctypes.data)
        for _ in range(3):
            R.nrf_fft_process(rfft, nb)
        reps_r = 30
        t0 = time.perf_counter()
        for _ in range(reps_r):
            R.nrf_fft_process(rfft, nb)
        t1 = time.perf_counter()
        for _ in range(reps_r):
            R.nrf_fft_process(rfft, nb); R.nut_buffer_free(R.nrf_fft_get_buffer(rfft))
        t2 = time.perf_counter()
        o = R.nrf_fft_get_buffer(rfft)
        ref_row = orc.nut_to_numpy(o)[:n]
        R.nut_buffer_free(o); R.nrf_fft_free(rfft); R.nut_buffer_free(nb)
        err = float(np.max(np.abs(last - ref_row)) / np.sqrt(np.mean(ref_row ** 2)))
        out["shapes"][f"{n}x{hist}"] = {"ours_process": ours_p, "ours_process_get_buffer": ours_pg,
                                       "reference_process": (t1 - t0) / reps_r * 1e6,
                                       "reference_process_get_buffer": (t2 - t1) / reps_r * 1e6, "row_rel_err": err}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
