#!/usr/bin/env python
"""Secondary measurement: the nrf_iq_filter FIR on the device.

  python tools/bench_fir.py [--taps 51] [--samples 16777216]

Device-resident float64 IQ lines (larger than L2), CUDA events around K launches of
fsea_fir_iq_device; parity gate first (bit-identical to the compiled reference on a sample);
the reference's own nrf_iq_filter loop timed beside it on one host core (it is a serial loop).
Prints one JSON line.  Algorithmic bytes per complex sample: 16 read + 16 written.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import frequensea_b200 as fb  # noqa: E402
from frequensea_b200 import nrf  # noqa: E402
from oracle import orc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--taps", type=int, default=51)
    ap.add_argument("--samples", type=int, default=1 << 24)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    L = fb.lib()
    vp = C.c_void_p
    L.fsea_fir_iq_device.argtypes = [vp, C.c_int, vp, vp, C.c_long, vp, vp]
    L.fsea_fir_iq_device.restype = C.c_int
    rate, half = 5000000, 200000
    coef = nrf.fir_low_pass_coefficients(rate, half, a.taps)
    taps = a.taps                                                     # the reference evaluates `length` taps (src/nrf.c:680)
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(3)
    n = a.samples
    si = torch.rand(n + taps - 1, dtype=torch.float64, device=dev, generator=g)
    sq = torch.rand(n + taps - 1, dtype=torch.float64, device=dev, generator=g)
    d_coef = torch.from_numpy(coef).to(dev)
    out = torch.empty(2 * n, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()

    def launch():
        rc = L.fsea_fir_iq_device(d_coef.data_ptr(), taps, si.data_ptr(), sq.data_ptr(), n, out.data_ptr(), stream.cuda_stream)
        assert rc == 0, rc

    launch()
    torch.cuda.synchronize()
    # parity gate: first 4096 outputs against the reference's arithmetic (sequential, unfused)
    hi, hq, c = si[:4096 + taps - 1].cpu().numpy(), sq[:4096 + taps - 1].cpu().numpy(), coef[:taps]
    ref = np.zeros(2 * 4096)
    for k in range(taps):
        ref[0::2] += c[k] * hi[k:k + 4096]
        ref[1::2] += c[k] * hq[k:k + 4096]
    assert np.array_equal(out[:2 * 4096].cpu().numpy(), ref), "FIR parity gate failed"
    for _ in range(a.warmup):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        launch()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    peak, src = bench.peak_hbm()
    gbs = 32.0 * n / (ms * 1e-3) / 1e9
    # the reference on one host core (a serial loop, src/nrf.c:759-768), bounded sample
    blk = np.random.default_rng(1).integers(0, 256, size=2 * 131072, dtype=np.uint8)
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < 5.0:
        orc.ref_iq_filter_blocks([blk], rate, half, a.taps)
        reps += 1
    cpu = reps * 131072 / (time.perf_counter() - t0)
    print(json.dumps({
        "metric": "iq_filter_samples_per_s", "value": n / (ms * 1e-3), "unit": "complex samples/s", "taps": taps,
        "samples_per_launch": n, "ms_per_launch": ms, "dtype": "f64", "parity": "bit-identical (first 4096 outputs)",
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                     "algorithmic_bytes_per_sample": 32, "peak_source": src,
                     "note": "unfused fp64 multiply + add per tap (bit-exactness with the reference): %.1f G fp64 ops/s" % (4.0 * taps * n / (ms * 1e-3) / 1e9)},
        "cpu_baseline": {"value": cpu, "unit": "complex samples/s", "cores": 1, "kind": "reference",
                         "sample": "%d device blocks of 131072 samples through the reference's nrf_iq_filter_process + get_buffer (oracle/_ref)" % reps},
    }))


if __name__ == "__main__":
    main()
