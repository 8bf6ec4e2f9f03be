"""TEST INFRASTRUCTURE: parity of the kernel SOURCE, executed on the CPU fiber machine
(tests/emu/libfsea_emu.so), against the oracle.  Run as a subprocess by
tests/test_emu_kernels.py with FSEA_TEST_LIB pointing at the emulated library.

    FSEA_TEST_LIB=tests/emu/libfsea_emu.so python tests/emu/check.py <n> [<n> ...]
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

from frequensea_b200 import binding as _binding  # noqa: E402
_binding._use_library(os.environ["FSEA_TEST_LIB"])      # the emulated build, explicitly (the product has no such switch)
import frequensea_b200 as fb  # noqa: E402
from conftest import synth_s8, synth_u8  # noqa: E402
from oracle import orc  # noqa: E402


def rel_err(got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    rms = np.sqrt(np.mean(ref ** 2, axis=-1, keepdims=True))
    return float(np.max(np.abs(got - ref) / rms))


def rows(data, n, hop, frames, fmt, win, out, scale=10.0, flags=0):
    return np.stack([orc.engine_frame(data[f * hop * 2:], n, fmt, win, out, scale, flags) for f in range(frames)])


def check(n, frames, arith=0):
    res = {"n": n}
    # (1) two's-complement input, magnitude, DC patch (nrf_fft semantics); more frames than clusters in flight
    data = synth_s8(n * frames, seed=n)
    plan = fb.Plan(n, in_format=fb.IN_S8, out_mode=fb.OUT_MAG, flags=fb.FLAG_DC_PATCH, arith=arith)
    res["kernel"] = plan.kernel_name
    got = plan.exec_host(data, frames)
    ref = rows(data, n, n, frames, orc.IN_S8, orc.WIN_NONE, orc.OUT_MAG, flags=orc.FLAG_DC_PATCH)
    res["mag_patch"] = rel_err(got, ref)
    res["patch_dup"] = bool(np.array_equal(got[:, n // 2], got[:, n // 2 - 1]))
    # (2) offset-binary input, uncentred DC kept, log-magnitude
    data = synth_u8(n * 3, seed=n + 1)
    plan = fb.Plan(n, in_format=fb.IN_U8, out_mode=fb.OUT_LOGMAG, arith=arith)
    got = plan.exec_host(data, 3).astype(np.float64)
    ref = rows(data, n, n, 3, orc.IN_U8, orc.WIN_NONE, orc.OUT_LOGMAG)
    lin = 10 ** (ref / 20)
    ok = lin >= 1e-3 * np.sqrt(np.mean(lin ** 2, axis=1, keepdims=True))
    res["logmag_db"] = float(np.max(np.abs(got - ref)[ok]))
    res["dc_bin_db"] = float(abs(got[0, n // 2] - ref[0, n // 2]))
    # (3) Hann + 50 % overlap + pixels (the sweep product)
    hop = n // 2
    data = synth_u8(2 * hop + n, seed=n + 2)
    plan = fb.Plan(n, hop=hop, in_format=fb.IN_U8, window=fb.WIN_HANN, out_mode=fb.OUT_PIX, pix_scale=5.0,
                   flags=fb.FLAG_DC_PATCH, arith=arith)
    got = plan.exec_host(data, 3)
    ref = rows(data, n, hop, 3, orc.IN_U8, orc.WIN_HANN, orc.OUT_PIX, 5.0, orc.FLAG_DC_PATCH)
    d = np.abs(got.astype(int) - ref.astype(int))
    res["pix_max"] = int(d.max())
    res["pix_frac"] = float((d > 0).mean())
    # (4) c/fft.c's scope row: bytes / 255, spectrum / N, intensity * log10(power + 1e-20)  (c/fft.c:50-51,181-197)
    data = synth_u8(n * 2, seed=n + 3)
    plan = fb.Plan(n, in_format=fb.IN_U8, out_mode=fb.OUT_LOGMAG, log_mult=7.0, power_gain=(256.0 / (255.0 * n)) ** 2, arith=arith)
    got = plan.exec_host(data, 2).astype(np.float64)
    ref = np.stack([orc.scope_row(data[f * 2 * n:], n, 7.0) for f in range(2)])
    lin = 10 ** (ref / 14)
    ok = lin >= 1e-3 * np.sqrt(np.mean(lin ** 2, axis=1, keepdims=True))
    res["scope_err"] = float(np.max(np.abs(got - ref)[ok]))
    return res


def check_file_feed():
    """fsea_exec_file == fsea_exec_host on the same bytes: offsets, overlap, short files."""
    import tempfile
    res = {"n": "file", "kernel": "fsea_exec_file"}
    data = synth_s8(40000, seed=11)
    with tempfile.NamedTemporaryFile(suffix=".raw", delete=False) as f:
        f.write(data.tobytes())
        path = f.name
    ok = True
    for n, hop, off_bytes, frames in ((1024, 1024, 0, 30), (1024, 512, 4096, 50), (256, 128, 16, 100), (16384, 16384, 0, 2)):
        plan = fb.Plan(n, hop=hop, in_format=fb.IN_S8, out_mode=fb.OUT_MAG, flags=fb.FLAG_DC_PATCH)
        a = plan.exec_file(path, frames, off_bytes)
        b = plan.exec_host(data[off_bytes:], frames)
        ok = ok and np.array_equal(a, b)
    plan = fb.Plan(1024, in_format=fb.IN_S8)
    for bad in (lambda: plan.exec_file(path, 40), lambda: plan.exec_file(path + ".missing", 1), lambda: plan.exec_file(path, 1, 8)):
        try:
            bad()
            ok = False
        except RuntimeError:
            pass
    os.unlink(path)
    res["identical"] = bool(ok)
    return res


def check_iq_filter():
    """nrf_iq_filter on top of fsea_fir_iq_kernel vs the compiled reference: bit-identical, block after block."""
    from frequensea_b200 import nrf
    res = {"n": "iq_filter", "kernel": "fsea_fir_iq_kernel"}
    rng = np.random.default_rng(5)
    worst = 0.0
    for taps, length in ((51, 1000), (50, 777), (5, 300), (3, 100), (2, 64), (1, 50), (97, 2500)):
        blocks = [rng.integers(0, 256, size=2 * length, dtype=np.uint8) for _ in range(3)]
        ref = orc.ref_iq_filter_blocks(blocks, 5000000, 200000, taps)
        flt = nrf.iq_filter_new(5000000, 200000, taps)
        for b, r in zip(blocks, ref):
            buf = nrf.nut_buffer_new_u8(length, 2, b)
            nrf.iq_filter_process(flt, buf)
            got = nrf.iq_filter_get_buffer(flt).numpy()
            worst = max(worst, float(np.max(np.abs(got - r))))
        nrf.iq_filter_free(flt)
        assert np.array_equal(nrf.fir_low_pass_coefficients(5000000, 200000, taps), orc.ref_fir_coefficients(5000000, 200000, taps))
    res["max_abs_diff"] = worst
    return res


if __name__ == "__main__":
    assert "emu" in fb.lib_path(), "this script checks the EMULATED library only"
    out = []
    if sys.argv[1:] == ["iq_filter"]:
        print(json.dumps([check_iq_filter()]))
        sys.exit(0)
    if sys.argv[1:] == ["file"]:
        print(json.dumps([check_file_feed()]))
        sys.exit(0)
    for a in sys.argv[1:]:                                   # N or N:arith
        n, _, ar = a.partition(":")
        n = int(n)
        out.append(check(n, 9 if n >= 8192 else 7, int(ar or 0)))
    print(json.dumps(out))
