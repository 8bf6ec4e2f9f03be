"""CPU suite: the oracle against (a) the golden values recorded in SURVEY.md
section 8(c), (b) numpy float64, (c) the compiled reference (oracle/_ref, when
present).  No GPU needed."""
import os

import numpy as np
import pytest

from oracle import orc
from conftest import BLOCK, synth_u8

# float64 golden values, block 0 / frame 0
SURVEY = {
    ("rf200big4", 1024): dict(sum=1547.983746054, max=103.947022303, argmax=871,
                              probe=[0.332605202459, 0.386267577875, 1.375575094689, 1.375575094689,
                                     0.214389733516, 0.325628273285], dc=714.180481248),
    ("rf200big4", 128): dict(sum=88.600227730, max=17.607009660, argmax=109,
                             probe=[0.228907583187, 0.237456154466, 0.438670479341, 0.438670479341,
                                    0.092331318054, 0.235718841052], dc=None),
    ("rf433", 1024): dict(sum=463.230317070, max=12.647431800, argmax=205,
                          probe=[0.177895274011, 0.127134335705, 0.495781753210, 0.495781753210,
                                 0.443210416628, 0.155556203745], dc=712.447206210),
    ("rf433", 128): dict(sum=16.717734564, max=1.276381726, argmax=26, probe=None, dc=None),
}


@pytest.mark.parametrize("key", list(SURVEY))
def test_oracle_matches_survey_golden(iq, key):
    name, n = key
    g = SURVEY[key]
    u8 = orc.device_unpack(iq[name][:BLOCK])
    mag = orc.nrf_fft_frame_u8(u8, n)
    assert abs(mag.sum() - g["sum"]) < 5e-9 * g["sum"]
    assert abs(mag.max() - g["max"]) < 5e-9 * g["max"]
    assert int(mag.argmax()) == g["argmax"]
    if g["probe"]:
        idx = [0, 1, n // 2 - 1, n // 2, n // 2 + 1, n - 1]
        np.testing.assert_allclose(mag[idx], g["probe"], rtol=0, atol=5e-12)
    if g["dc"]:
        out = orc.engine_frame(u8, n, orc.IN_U8, orc.WIN_NONE, orc.OUT_MAG)   # no patch: DC visible
        assert abs(out[n // 2] - g["dc"]) < 1e-8


def test_oracle_pixels_match_survey_golden(iq):
    raw = iq["rf200big4"]
    pix, _ = orc.batch_row(raw, 1024, 10.0, False)      # c/fft-batch.c
    assert int(pix.sum()) == 23772 and list(pix[510:515]) == [42, 27, 255, 0, 0]
    assert int(pix.max()) == 255 and int(pix.argmax()) == 461
    pix, _ = orc.batch_row(raw, 256, 5.0, True)         # c/fft-batch-broad.c
    assert int(pix.sum()) == 1371 and list(pix[126:131]) == [0, 0, 0, 0, 4]
    assert int(pix.max()) == 147 and int(pix.argmax()) == 218


def test_oracle_matches_numpy_golden(iq, golden):
    for name in ("rf433", "rf200big"):
        raw = iq["rf433" if name == "rf433" else "rf200big4"]
        u8 = orc.device_unpack(raw[:BLOCK])
        for n in (128, 1024):
            frames = u8.reshape(-1, 2 * n)
            got = np.stack([orc.nrf_fft_frame_u8(f, n) for f in frames[:8]])
            np.testing.assert_allclose(got, golden[f"np_mag_{name}_n{n}_first8"], rtol=0, atol=2e-12 * n)
            sums = np.array([orc.nrf_fft_frame_u8(f, n).sum() for f in frames])
            np.testing.assert_allclose(sums, golden[f"np_magsum_{name}_n{n}"], rtol=1e-12)
    raw = iq["rf200big4"]
    for b in range(4):
        assert np.array_equal(orc.batch_row(raw[b * BLOCK:], 1024, 10.0, False)[0], golden["np_pix_batch_n1024"][b])
        assert np.array_equal(orc.batch_row(raw[b * BLOCK:], 256, 5.0, True)[0], golden["np_pix_broad_n256"][b])


def test_engine_definition_equals_reference_statements(iq):
    """orc.engine_frame (the general definition the GPU is tested against) must
    reduce to the reference-specific restatements."""
    raw = iq["rf200big4"][:BLOCK]
    u8 = orc.device_unpack(raw)
    for n in (16, 128, 1024, 4096):
        a = orc.nrf_fft_frame_u8(u8, n)
        b = orc.engine_frame(u8, n, orc.IN_U8, orc.WIN_NONE, orc.OUT_MAG, flags=orc.FLAG_DC_PATCH)
        c = orc.engine_frame(raw, n, orc.IN_S8, orc.WIN_NONE, orc.OUT_MAG, flags=orc.FLAG_DC_PATCH)
        assert np.array_equal(a, b) and np.array_equal(a, c)
    p1 = orc.batch_row(raw, 1024, 10.0, False)[0]
    p2 = orc.engine_frame(raw, 1024, orc.IN_S8, orc.WIN_NONE, orc.OUT_PIX, pix_scale=10.0)
    assert np.array_equal(p1, p2)
    p1 = orc.batch_row(raw, 256, 5.0, True)[0]
    p2 = orc.engine_frame(raw, 256, orc.IN_S8, orc.WIN_NONE, orc.OUT_PIX, pix_scale=5.0, flags=orc.FLAG_DC_PATCH)
    assert np.array_equal(p1, p2)


def test_dft_against_naive_and_numpy():
    rng = np.random.default_rng(3)
    for n in (1, 2, 8, 64, 256, 24):
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        ref = np.fft.fft(x)
        assert np.abs(orc.dft_forward(x) - ref).max() < 1e-12 * max(1, n)
        assert np.abs(orc.dft_forward(x, naive=True) - ref).max() < 1e-12 * max(1, n)
    x = rng.standard_normal(65536) + 1j * rng.standard_normal(65536)
    assert np.abs(orc.dft_forward(x) - np.fft.fft(x)).max() < 1e-9


def test_hann_definition():
    n = 256
    u8 = synth_u8(n)
    got = orc.engine_frame(u8, n, orc.IN_U8, orc.WIN_HANN, orc.OUT_LOGMAG)
    w = 0.5 * (1 - np.cos(2 * np.pi * np.arange(n) / n))
    x = ((u8[0::2].astype(float) - 128) + 1j * (u8[1::2].astype(float) - 128)) / 256 * w * (1 - 2 * (np.arange(n) % 2))
    ref = 10 * np.log10(np.abs(np.fft.fft(x)) ** 2 + 1e-20)
    np.testing.assert_allclose(got, ref, atol=1e-9)


def test_history_and_shift_against_reference_semantics(golden):
    h = np.zeros((4, 8))
    for r in range(6):
        orc.history_push(h, np.full(8, r + 1.0))
    assert [row[0] for row in h] == [6.0, 5.0, 4.0, 3.0]       # row 0 newest (src/nrf.c:617)
    if "ref_shift_input" in golden:
        for d in (3.7, -9.0, 0.9, 1e9):
            img = golden["ref_shift_input"].copy()
            orc.fft_shift(img, d)
            assert np.array_equal(img, golden[f"ref_shift_d{d}"])


def test_stitch_definition():
    """c/fft-stitch-broad.c:28-36,66-88: max-blend at i * WIDTH_STEP."""
    rng = np.random.default_rng(2)
    imgs = rng.integers(0, 256, (6, 9, 32), dtype=np.uint8)
    for step in (32, 12, 5, 50):
        ref = np.zeros((9, 32 + 5 * step), np.uint8)
        for i in range(6):
            ref[:, i * step:i * step + 32] = np.maximum(ref[:, i * step:i * step + 32], imgs[i])
        assert np.array_equal(orc.stitch(imgs, step), ref)


def test_scope_row_definition(iq):
    raw = iq["rf433"]
    n = 512
    got = orc.scope_row(raw, n, 10.0)
    x = (raw[0:2 * n:2].astype(float) + 1j * raw[1:2 * n:2].astype(float)) / 255.0
    X = np.fft.fftshift(np.fft.fft(x)) / n
    np.testing.assert_allclose(got, 10 * np.log10(np.abs(X) ** 2 + 1e-20), atol=1e-9)


# ---- against the compiled reference itself ---------------------------------
needs_ref = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref/libnrf_ref.so not built (needs /root/reference)")


@needs_ref
def test_oracle_equals_compiled_reference(iq, golden):
    raw = iq["rf200big4"]
    blocks = [orc.device_unpack(raw[b * BLOCK:(b + 1) * BLOCK]) for b in range(4)]
    for n in (128, 1024):
        hist = orc.ref_nrf_fft_rows(None, n, 6, blocks)
        assert np.array_equal(hist, golden[f"ref_nrf_fft_hist_n{n}_h6"])
        for b in range(4):                                   # newest first
            mine = orc.nrf_fft_frame_u8(blocks[b], n)
            np.testing.assert_allclose(hist[3 - b], mine, rtol=0, atol=1e-12)
        assert not hist[4:].any()


@needs_ref
def test_reference_f64_path_and_shifter(golden):
    R = orc.ref()
    rng = np.random.default_rng(5)
    x = rng.standard_normal(2 * 1024)
    nb = R.nut_buffer_new_f64(1024, 2, x.ctypes.data)
    fft = R.nrf_fft_new(1024, 1)
    R.nrf_fft_process(fft, nb)
    out = R.nrf_fft_get_buffer(fft)
    got = orc.nut_to_numpy(out)
    np.testing.assert_allclose(got, orc.nrf_fft_frame_f64(x, 1024), rtol=0, atol=1e-11)
    for b in (nb, out):
        R.nut_buffer_free(b)
    R.nrf_fft_free(fft)
