"""CPU suite: the C-ABI library loads, exports every symbol the headers declare,
rejects bad arguments, and refuses to compute without a CUDA device (no CPU
fallback).  The host-side container (nut_*) is checked against the compiled
reference where that exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b((?:fsea|nrf|nut)_[a-z0-9_]+)\s*\(", src))
    return {n for n in names if not n.endswith("_fn")}


def test_library_exports_every_declared_symbol(fb):
    L = fb.lib()
    missing = []
    for h in ("fsea.h", "nrf.h", "nut.h"):
        for name in sorted(_declared(h)):
            if not hasattr(L, name):
                missing.append(f"{h}:{name}")
    assert not missing, missing


def test_error_codes_and_argument_checks(fb):
    L = fb.lib()
    assert L.fsea_strerror(0) == b"ok"
    assert b"no CPU fallback" in L.fsea_strerror(-3)
    with pytest.raises(fb.FseaError) as e:
        fb.Plan(1000)                         # not a power of two
    assert e.value.code == -2
    with pytest.raises(fb.FseaError) as e:
        fb.Plan(8)                            # below the supported range
    assert e.value.code == -2
    with pytest.raises(fb.FseaError) as e:
        fb.Plan(1024, in_format=9)
    assert e.value.code == -1


def test_no_cpu_fallback(fb):
    if fb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(fb.FseaError) as e:
        fb.Plan(1024)
    assert e.value.code == -3                 # FSEA_ERR_NO_DEVICE


def test_config_struct_layout(fb):
    from frequensea_b200.binding import FseaConfig
    assert C.sizeof(FseaConfig) == 64
    cfg = FseaConfig()
    fb.lib().fsea_config_default(C.byref(cfg), 4096)
    assert (cfg.fft_size, cfg.hop, cfg.in_format, cfg.out_mode, cfg.pix_scale) == (4096, 4096, 0, 0, 10.0)


def test_nut_buffer_semantics(fb):
    nrf = fb.nrf
    data = np.arange(24, dtype=np.uint8)
    b = nrf.nut_buffer_new_u8(12, 2, data)
    assert (b.type, b.length, b.channels, b.size_bytes) == (1, 12, 2, 24)
    assert np.array_equal(b.numpy(), data)
    f = nrf.nut_buffer_convert(b, nrf.NUT_BUFFER_F64)
    assert f.size_bytes == 24 * 8 and np.array_equal(f.numpy(), data / 256.0)       # src/nut.c:129-135
    back = nrf.nut_buffer_convert(f, nrf.NUT_BUFFER_U8)
    assert np.array_equal(back.numpy(), data)
    r = nrf.nut_buffer_reduce(b, 0.5)
    assert r.length == 6 and np.array_equal(r.numpy(), data[:12])
    c = nrf.nut_buffer_clip(b, 4, 3)                                                 # offset in elements
    assert c.length == 3 and np.array_equal(c.numpy(), data[4:10])
    c2 = nrf.nut_buffer_clip(b, 4, -1)
    assert c2.length == 8
    a = nrf.nut_buffer_new_u8(2, 2, np.array([9, 8, 7, 6], dtype=np.uint8))
    nrf.nut_buffer_append(a, b)
    assert a.length == 14 and a.size_bytes == 28 and np.array_equal(a.numpy()[:4], [9, 8, 7, 6])
    z = nrf.nut_buffer_new_f64(3, 1)
    assert np.array_equal(z.numpy(), np.zeros(3))


@pytest.mark.skipif(not orc.have_ref(), reason="compiled reference not available")
def test_nut_buffer_equals_compiled_reference(fb):
    """Same calls into the reference's own src/nut.c and into libfsea.so."""
    R, L = orc.ref(), fb.lib()
    fb.nrf._L()
    rng = np.random.default_rng(0)
    data = rng.integers(0, 256, 64, dtype=np.uint8)
    for lib_, NB in ((R, orc.NutBuffer), (L, fb.nrf.NutBufferStruct)):
        pass
    def run(lib_):
        b = lib_.nut_buffer_new_u8(32, 2, data.ctypes.data)
        f = lib_.nut_buffer_convert(b, 2)
        r = lib_.nut_buffer_reduce(f, 0.37)
        c = lib_.nut_buffer_clip(b, 5, 7)
        lib_.nut_buffer_append(b, c)
        outs = []
        for x in (b, f, r, c):
            s = x.contents
            outs.append((s.type, s.length, s.channels, s.size_bytes, orc.nut_to_numpy(x)))
        return outs
    for a, b in zip(run(R), run(L)):
        assert a[:4] == b[:4] and np.array_equal(a[4], b[4])


def test_freq_shifter_host_matches_reference_golden(fb, iq, golden):
    if "ref_shifter_out_first4096" not in golden:
        pytest.skip("golden without reference arrays")
    nrf = fb.nrf
    u8 = orc.device_unpack(iq["rf200big4"][:262144])
    sh = nrf.freq_shifter_new(-10000, 5000000)
    b = nrf.nut_buffer_new_u8(131072, 2, u8)
    nrf.freq_shifter_process(sh, b)
    out = nrf.freq_shifter_get_buffer(sh)
    assert [out.length, out.channels] == list(golden["ref_shifter_len_channels"])
    np.testing.assert_allclose(out.numpy()[:4096], golden["ref_shifter_out_first4096"], rtol=0, atol=1e-15)
    nrf.freq_shifter_free(sh)


def test_device_file_replay_semantics(fb, iq, tmp_path):
    """nrf_device dummy path: (b+128)%256 unpack, block stepping, pause (src/nrf.c:95-110,341-357)."""
    nrf = fb.nrf
    p = tmp_path / "cap.raw"
    iq["rf200big4"].tofile(p)
    dev = nrf.device_new(97.0, str(p))
    nrf.device_set_paused(dev, True)
    import time
    time.sleep(0.1)
    seen = set()
    for step in range(4):
        time.sleep(0.06)                       # > one 60 Hz period: the thread republishes the block
        got = nrf.device_get_samples_buffer(dev)
        assert (got.type, got.length, got.channels, got.size_bytes) == (1, 131072, 2, 262144)
        arr = got.numpy()
        matches = [b for b in range(4) if np.array_equal(arr, orc.device_unpack(iq["rf200big4"][b * 262144:(b + 1) * 262144]))]
        assert len(matches) == 1
        seen.add(matches[0])
        nrf.device_step(dev)
    assert len(seen) >= 3
    assert nrf.device_set_frequency(dev, 123.4) == 123.4
    nrf.device_free(dev)


def test_device_short_capture_file_is_padded_to_one_block(fb, iq, tmp_path):
    """rfdata/rf-433.000-short.raw is 172144 bytes, less than one 262144-byte block: the reference computes
    dummy_block_length = size / 262144 = 0 and reads out of bounds (src/nrf.c:268,162-166).  Here the file is
    padded with silence to one block: the bytes that exist are published ((b+128)%256), the rest is 0x80."""
    nrf = fb.nrf
    short = iq["rf433"][:172144]
    p = tmp_path / "rf-433.000-short.raw"
    short.tofile(p)
    dev = nrf.device_new(433.0, str(p))
    nrf.device_set_paused(dev, True)
    import time
    for _ in range(100):                                   # the 60 Hz producer thread publishes the block
        got = nrf.device_get_samples_buffer(dev).numpy()
        if got[:64].any():
            break
        time.sleep(0.02)
    assert got.size == 262144
    assert np.array_equal(got[:172144], orc.device_unpack(short))
    assert np.all(got[172144:] == 0x80)
    nrf.device_step(dev)                                   # one block only: stepping wraps to the same block
    time.sleep(0.05)
    assert np.array_equal(nrf.device_get_samples_buffer(dev).numpy()[:172144], orc.device_unpack(short))
    nrf.device_free(dev)


def test_mgpu_library_exports_every_declared_symbol(fb):
    """include/fsea_mgpu.h <-> libfsea_mgpu.so (the one-process NCCL driver): load + symbols, no compute."""
    path = os.path.join(ROOT, "frequensea_b200", "libfsea_mgpu.so")
    if not os.path.exists(path):
        fb.build()
    fb.lib()                                               # libfsea.so first (DT_NEEDED by rpath too)
    L = C.CDLL(path)
    for name in sorted(_declared("fsea_mgpu.h")):
        assert hasattr(L, name), name


def test_fir_coefficients_match_reference_golden(fb, golden):
    """nrf_fir_get_low_pass_coefficients (src/nrf.c:654-676): host code, no device needed -- bit-identical to
    the coefficients recorded from the compiled reference, odd and even lengths."""
    from frequensea_b200 import nrf
    for taps, half in ((51, 200000), (97, 60000), (50, 100000)):
        got = nrf.fir_low_pass_coefficients(5000000, half, taps)
        ref = golden[f"ref_fir_coef_t{taps}_f{half}"]
        assert got.shape == ref.shape and np.array_equal(got, ref)
        assert abs(got.sum() - 1.0) < 1e-12                       # unit DC gain
