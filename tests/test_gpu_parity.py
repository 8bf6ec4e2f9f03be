"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the
oracle on the same inputs.

Tolerances (north star: spectra within 1e-5 relative of a float64 DFT):
  MAG     |y - y_ref| <= 1e-5 * rms_k(y_ref) + 1e-6 * |y_ref|   every bin of every frame (fp32 kernels)
          (atol/rtol form.  The rtol term only matters on tone bins far above the frame's rms: fp32
          butterflies leave sigma ~ 1.1e-7 relative per bin at N = 8192 -- DESIGN.md 4 -- so a bin 52 x rms,
          as in the cfg3 generator, cannot be promised to 1e-5 * rms = 1.6 ulp; it is held to 1e-6 of itself.)
  LOGMAG  |dB - dB_ref| <= 2e-3 dB wherever |X_ref| >= 1e-3 * rms (log of a cancelling bin is ill-conditioned)
  PIX     |pix - pix_ref| <= 1 and fewer than 2e-4 of the pixels differ (trunc() of a float; measured 8e-6 ... 3e-5)
  F64     max |y - y_ref| <= 1e-11 * rms
"""
import os

import numpy as np
import pytest

from oracle import orc
from conftest import BLOCK, synth_s8, synth_u8

pytestmark = pytest.mark.gpu

MAG_TOL = 1e-5
# u8 pixels are trunc() of a float: the fp32 spectrum's ~1e-6 relative error moves S * dB by ~4e-5 S / 10, so about that
# fraction of pixels lands on the other side of an integer (measured on 10^8 pixels of the cfg4 job and in every test
# here: 8e-6 ... 3e-5 differing, always by exactly 1).  Bound: 1 LSB and 2e-4 of the pixels (a handful in a small test).
PIX_DIFF_FRAC = 2e-4


def pix_close(d):
    d = np.asarray(d)
    return bool(d.max() <= 1 and (d > 0).sum() <= max(3, PIX_DIFF_FRAC * d.size))


def rel_err(got, ref):
    """max over bins of |got - ref| / (rms + 0.1 |ref|): `< 1e-5` is |d| <= 1e-5 rms + 1e-6 |ref|."""
    ref = np.asarray(ref, dtype=np.float64)
    rms = np.sqrt(np.mean(ref ** 2, axis=-1, keepdims=True))
    return float(np.max(np.abs(got - ref) / (rms + 0.1 * np.abs(ref))))


def oracle_rows(data, n, hop, frames, fmt, win, out, scale=10.0, flags=0):
    per = 2
    rows = []
    for f in range(frames):
        rows.append(orc.engine_frame(data[f * hop * per:], n, fmt, win, out, scale, flags))
    return np.stack(rows)


@pytest.fixture(scope="module")
def fbg(fb):
    assert fb.device_count() > 0, "-m gpu tests need a CUDA device (the product has no CPU path)"
    return fb


@pytest.mark.parametrize("arith", [1, 2])
@pytest.mark.parametrize("n", [128, 1024])
def test_golden_fixture_frames(fbg, iq, golden, n, arith):
    """cfg2: every frame of the reference's capture block, nrf_fft semantics."""
    fb = fbg
    for name, raw in (("rf433", iq["rf433"]), ("rf200big", iq["rf200big4"][:BLOCK])):
        frames = BLOCK // (2 * n)
        plan = fb.Plan(n, in_format=fb.IN_S8, out_mode=fb.OUT_MAG, flags=fb.FLAG_DC_PATCH, arith=arith)
        got = plan.exec_host(raw, frames)
        first8 = golden[f"np_mag_{name}_n{n}_first8"]
        assert rel_err(got[:8], first8) < MAG_TOL
        sums = golden[f"np_magsum_{name}_n{n}"]
        np.testing.assert_allclose(got.astype(np.float64).sum(axis=1), sums, rtol=2e-6)
        ref = oracle_rows(raw, n, n, frames, orc.IN_S8, orc.WIN_NONE, orc.OUT_MAG, flags=orc.FLAG_DC_PATCH)
        assert rel_err(got, ref) < MAG_TOL
        # the reference's DC quirk: bin N/2 duplicates bin N/2-1 exactly
        assert np.array_equal(got[:, n // 2], got[:, n // 2 - 1])
        # and without the patch the uncentred DC term is there (|X[N/2]| ~ N/2*sqrt(2))
        plan2 = fb.Plan(n, in_format=fb.IN_S8, out_mode=fb.OUT_MAG, arith=arith)
        got2 = plan2.exec_host(raw, 1)
        assert abs(got2[0, n // 2] - float(golden[f"np_dcbin_{name}_n{n}"])) < 1e-4 * n


@pytest.mark.parametrize("n", [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("arith", [1, 2])
def test_all_sizes_mag(fbg, n, arith):
    fb = fbg
    frames = 7                                   # odd: exercises the half-empty frame pair
    data = synth_u8(n * frames, seed=n)
    plan = fb.Plan(n, in_format=fb.IN_U8, out_mode=fb.OUT_MAG, flags=fb.FLAG_DC_PATCH, arith=arith)
    got = plan.exec_host(data, frames)
    ref = oracle_rows(data, n, n, frames, orc.IN_U8, orc.WIN_NONE, orc.OUT_MAG, flags=orc.FLAG_DC_PATCH)
    assert rel_err(got, ref) < MAG_TOL, plan.kernel_name
    plan = fb.Plan(n, in_format=fb.IN_U8, out_mode=fb.OUT_MAG, arith=arith)    # uncentred DC kept
    got = plan.exec_host(data, frames)
    ref = oracle_rows(data, n, n, frames, orc.IN_U8, orc.WIN_NONE, orc.OUT_MAG)
    k = np.arange(n) != n // 2
    assert rel_err(got[:, k], ref[:, k]) < MAG_TOL
    np.testing.assert_allclose(got[:, n // 2], ref[:, n // 2], rtol=2e-6)


@pytest.mark.parametrize("n,hop", [(256, 128), (1024, 512), (1024, 1000), (8192, 4096), (64, 8)])
def test_hann_overlap_logmag(fbg, n, hop):
    """cfg3 shape: Hann window, overlapping frames, log-magnitude."""
    fb = fbg
    frames = 9
    data = synth_u8((frames - 1) * hop + n, seed=hop)
    plan = fb.Plan(n, hop=hop, in_format=fb.IN_U8, window=fb.WIN_HANN, out_mode=fb.OUT_LOGMAG)
    got = plan.exec_host(data, frames).astype(np.float64)
    ref = oracle_rows(data, n, hop, frames, orc.IN_U8, orc.WIN_HANN, orc.OUT_LOGMAG)
    lin = 10 ** (ref / 20)
    ok = lin >= 1e-3 * np.sqrt(np.mean(lin ** 2, axis=1, keepdims=True))
    assert ok.mean() > 0.95
    assert np.max(np.abs(got - ref)[ok]) < 2e-3
    plan = fb.Plan(n, hop=hop, in_format=fb.IN_U8, window=fb.WIN_HANN, out_mode=fb.OUT_MAG)
    got = plan.exec_host(data, frames)
    ref = oracle_rows(data, n, hop, frames, orc.IN_U8, orc.WIN_HANN, orc.OUT_MAG)
    assert rel_err(got, ref) < MAG_TOL


@pytest.mark.parametrize("arith", [0, 3])
@pytest.mark.parametrize("n", [2048, 4096, 8192])
def test_frame_per_unit_variants(fbg, n, arith):
    """One frame per unit (packed lanes = even / odd samples, radix-2 across the lanes in the output
    stage) and whatever FSEA_ARITH_AUTO picks for these lengths (N = 4096 without a window: 64 x 64
    with 64 points per thread, fsea_k64.cuh): all output modes, odd frame counts."""
    fb = fbg
    frames = 5
    data = synth_u8(n * frames, seed=n + 3)
    for flags in (fb.FLAG_DC_PATCH, 0):
        plan = fb.Plan(n, in_format=fb.IN_U8, out_mode=fb.OUT_MAG, flags=flags, arith=arith)
        got = plan.exec_host(data, frames)
        ref = oracle_rows(data, n, n, frames, orc.IN_U8, orc.WIN_NONE, orc.OUT_MAG, flags=flags)
        k = np.arange(n) != n // 2 if not flags else np.ones(n, bool)
        assert rel_err(got[:, k], ref[:, k]) < MAG_TOL, plan.kernel_name
        if flags:
            assert np.array_equal(got[:, n // 2], got[:, n // 2 - 1])
        else:
            assert abs(got[0, n // 2] - ref[0, n // 2]) < 1e-4 * n
    data = synth_s8(n * frames, seed=n + 4)
    plan = fb.Plan(n, in_format=fb.IN_S8, out_mode=fb.OUT_LOGMAG, arith=arith)
    got = plan.exec_host(data, frames).astype(np.float64)
    ref = oracle_rows(data, n, n, frames, orc.IN_S8, orc.WIN_NONE, orc.OUT_LOGMAG)
    lin = 10 ** (ref / 20)
    ok = lin >= 1e-3 * np.sqrt(np.mean(lin ** 2, axis=1, keepdims=True))
    assert np.max(np.abs(got - ref)[ok]) < 2e-3, plan.kernel_name
    hop = n // 2
    data = synth_u8((frames - 1) * hop + n, seed=n + 5)
    plan = fb.Plan(n, hop=hop, in_format=fb.IN_U8, window=fb.WIN_HANN, out_mode=fb.OUT_PIX, pix_scale=5.0,
                   flags=fb.FLAG_DC_PATCH, arith=arith)
    got = plan.exec_host(data, frames)
    ref = oracle_rows(data, n, hop, frames, orc.IN_U8, orc.WIN_HANN, orc.OUT_PIX, 5.0, orc.FLAG_DC_PATCH)
    d = np.abs(got.astype(int) - ref.astype(int))
    assert pix_close(d), plan.kernel_name


@pytest.mark.parametrize("n", [16384, 32768, 65536])
def test_four_step_sizes(fbg, n):
    """cfg4 / cfg5 sizes: N = N1 x N2 through the L2-resident intermediate."""
    fb = fbg
    frames = 3
    data = synth_s8(n * frames, seed=n)
    plan = fb.Plan(n, in_format=fb.IN_S8, out_mode=fb.OUT_MAG, flags=fb.FLAG_DC_PATCH)
    got = plan.exec_host(data, frames)
    ref = oracle_rows(data, n, n, frames, orc.IN_S8, orc.WIN_NONE, orc.OUT_MAG, flags=orc.FLAG_DC_PATCH)
    assert rel_err(got, ref) < MAG_TOL, plan.kernel_name
    assert np.array_equal(got[:, n // 2], got[:, n // 2 - 1])
    # uncentred DC kept (no patch), log-magnitude as cfg5 asks
    plan = fb.Plan(n, in_format=fb.IN_S8, out_mode=fb.OUT_LOGMAG)
    got = plan.exec_host(data, frames).astype(np.float64)
    ref = oracle_rows(data, n, n, frames, orc.IN_S8, orc.WIN_NONE, orc.OUT_LOGMAG)
    lin = 10 ** (ref / 20)
    ok = lin >= 1e-3 * np.sqrt(np.mean(lin ** 2, axis=1, keepdims=True))
    assert np.max(np.abs(got - ref)[ok]) < 2e-3
    assert abs(got[0, n // 2] - ref[0, n // 2]) < 2e-3          # the analytic DC bin
    # Hann + 50 % overlap + pixels (sweep product), odd frame count
    hop = n // 2
    data = synth_u8((frames - 1) * hop + n, seed=n + 1)
    plan = fb.Plan(n, hop=hop, in_format=fb.IN_U8, window=fb.WIN_HANN, out_mode=fb.OUT_PIX, pix_scale=5.0, flags=fb.FLAG_DC_PATCH)
    got = plan.exec_host(data, frames)
    ref = oracle_rows(data, n, hop, frames, orc.IN_U8, orc.WIN_HANN, orc.OUT_PIX, 5.0, orc.FLAG_DC_PATCH)
    d = np.abs(got.astype(int) - ref.astype(int))
    assert pix_close(d)


def test_four_step_many_frames(fbg):
    """More frames than one L2-sized chunk (chunk = 32 MiB of intermediate = 64 frames at N = 65536)."""
    import torch
    fb = fbg
    n, frames = 65536, 150
    g = torch.Generator(device="cuda").manual_seed(3)
    d_in = torch.randint(0, 256, (frames * n * 2,), dtype=torch.uint8, device="cuda", generator=g)
    d_out = torch.empty((frames, n), dtype=torch.float32, device="cuda")
    plan = fb.Plan(n, in_format=fb.IN_S8, out_mode=fb.OUT_MAG)
    plan.exec_device(d_in.data_ptr(), d_out.data_ptr(), frames, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    x = (d_in.view(frames, n, 2).to(torch.int16) ^ 0x80).to(torch.float64) / 256.0
    energy_t = (x ** 2).sum(dim=(1, 2)) * n
    energy_f = (d_out.to(torch.float64) ** 2).sum(dim=1)
    assert torch.max(torch.abs(energy_f / energy_t - 1)).item() < 1e-5          # Parseval
    h_in = d_in.cpu().numpy()
    for f in (0, 63, 64, 149):
        ref = orc.engine_frame(h_in[f * 2 * n:], n, orc.IN_S8, orc.WIN_NONE, orc.OUT_MAG)
        k = np.arange(n) != n // 2
        assert rel_err(d_out[f].cpu().numpy()[k], ref[k]) < MAG_TOL


def test_sweep_pixels_match_reference_tools(fbg, iq, golden):
    """cfg1 / cfg4 semantics: c/fft-batch.c (S=10, no patch) and c/fft-batch-broad.c (S=5, DC patch)."""
    fb = fbg
    raw = iq["rf200big4"]
    for n, scale, flags, key in ((1024, 10.0, 0, "np_pix_batch_n1024"), (256, 5.0, fb.FLAG_DC_PATCH, "np_pix_broad_n256")):
        plan = fb.Plan(n, hop=BLOCK // 2, in_format=fb.IN_S8, out_mode=fb.OUT_PIX, pix_scale=scale, flags=flags)
        got = plan.exec_host(raw, 4)                       # one row per 262144-byte block, as the tools do
        ref = golden[key]
        d = np.abs(got.astype(int) - ref.astype(int))
        assert pix_close(d)
        if n == 1024:
            assert got[0, 512] == 255                      # SURVEY golden: DC pixel saturates without the patch
        # all frames of block 0 against the oracle
        frames = BLOCK // (2 * n)
        plan = fb.Plan(n, in_format=fb.IN_S8, out_mode=fb.OUT_PIX, pix_scale=scale, flags=flags)
        got = plan.exec_host(raw, frames)
        ref = oracle_rows(raw, n, n, frames, orc.IN_S8, orc.WIN_NONE, orc.OUT_PIX, scale, flags)
        d = np.abs(got.astype(int) - ref.astype(int))
        assert pix_close(d)


@pytest.mark.parametrize("n", [16, 256, 1024, 8192, 16384, 32768, 65536])
def test_float64_arithmetic_pixels_exact(fbg, n):
    """FSEA_ARITH_F64: 8-bit samples widened and windowed in double on the device, transformed in fp64 -- the
    arithmetic of the reference (src/nrf.c:599-615, c/fft-batch.c:61-92).  Pixels are the oracle's BYTE FOR BYTE
    (both tools' scales, with and without the Hann window / overlap), float rows agree to 1e-11 of the rms."""
    fb = fbg
    frames = 5 if n <= 8192 else 3
    data = synth_s8(n * frames, seed=n + 41)
    for scale, flags in ((10.0, 0), (5.0, fb.FLAG_DC_PATCH)):                 # c/fft-batch.c, c/fft-batch-broad.c
        plan = fb.Plan(n, in_format=fb.IN_S8, out_mode=fb.OUT_PIX, pix_scale=scale, flags=flags, arith=fb.ARITH_F64)
        got = plan.exec_host(data, frames)
        ref = oracle_rows(data, n, n, frames, orc.IN_S8, orc.WIN_NONE, orc.OUT_PIX, scale, orc.FLAG_DC_PATCH if flags else 0)
        assert got.dtype == np.uint8 and np.array_equal(got, ref), plan.kernel_name
    hop = n // 2
    data = synth_u8((frames - 1) * hop + n, seed=n + 42)
    plan = fb.Plan(n, hop=hop, in_format=fb.IN_U8, window=fb.WIN_HANN, out_mode=fb.OUT_PIX, pix_scale=5.0,
                   flags=fb.FLAG_DC_PATCH, arith=fb.ARITH_F64)
    got = plan.exec_host(data, frames)
    ref = oracle_rows(data, n, hop, frames, orc.IN_U8, orc.WIN_HANN, orc.OUT_PIX, 5.0, orc.FLAG_DC_PATCH)
    assert np.array_equal(got, ref), plan.kernel_name
    plan = fb.Plan(n, hop=hop, in_format=fb.IN_U8, window=fb.WIN_HANN, out_mode=fb.OUT_MAG, arith=fb.ARITH_F64)
    got = plan.exec_host(data, frames)
    ref = oracle_rows(data, n, hop, frames, orc.IN_U8, orc.WIN_HANN, orc.OUT_MAG)
    assert got.dtype == np.float64
    assert np.max(np.abs(got - ref)) <= 1e-11 * np.sqrt(np.mean(ref ** 2))
    with pytest.raises(fb.FseaError):
        fb.Plan(n, in_format=fb.IN_F32C, arith=fb.ARITH_F64)                 # float samples pick their arithmetic themselves


def test_float64_arithmetic_reference_tool_rows(fbg, iq, golden):
    """The rows c/fft-batch.c and c/fft-batch-broad.c write for the committed capture blocks (golden: computed by
    the compiled reference), byte for byte."""
    fb = fbg
    raw = iq["rf200big4"]
    for n, scale, flags, key in ((1024, 10.0, 0, "np_pix_batch_n1024"), (256, 5.0, fb.FLAG_DC_PATCH, "np_pix_broad_n256")):
        plan = fb.Plan(n, hop=BLOCK // 2, in_format=fb.IN_S8, out_mode=fb.OUT_PIX, pix_scale=scale, flags=flags, arith=fb.ARITH_F64)
        got = plan.exec_host(raw, 4)
        assert np.array_equal(got, golden[key]), key


@pytest.mark.parametrize("n", [512, 1024, 16384])
def test_scope_row_matches_c_fft(fbg, iq, n):
    """c/fft.c:45-56,181-197 (the 512-point scope): raw bytes / 255 with no offset conversion,
    half-swap fftshift, spectrum / N, intensity * log10(power + 1e-20) -- as one LOGMAG plan with
    log_mult = intensity, power_gain = (256 / (255 N))^2."""
    fb = fbg
    raw = iq["rf200big4"]
    frames, intensity = 6, 10.0
    plan = fb.Plan(n, in_format=fb.IN_U8, out_mode=fb.OUT_LOGMAG, log_mult=intensity,
                   power_gain=(256.0 / (255.0 * n)) ** 2)
    got = plan.exec_host(raw, frames).astype(np.float64)
    ref = np.stack([orc.scope_row(raw[f * 2 * n:], n, intensity) for f in range(frames)])
    lin = 10 ** (ref / (2 * intensity))
    ok = lin >= 1e-3 * np.sqrt(np.mean(lin ** 2, axis=1, keepdims=True))
    assert ok.mean() > 0.95
    assert np.max(np.abs(got - ref)[ok]) < 2e-3, plan.kernel_name


def test_glsl_log_power_consumer_constants(fbg, iq):
    """the fragment shader of lua/fft-shifted.lua:19-35 turns the two texture channels into
    v = 10 ln(r^2 + g^2 + 1e-20) * 0.02.  With the spectrum's re / im as the channels that is LOGMAG with
    log_mult = 0.2 ln 10 and power_gain = 1: checked at exactly those constants against float64."""
    fb = fbg
    raw = iq["rf433"]
    n, frames = 1024, 16
    plan = fb.Plan(n, in_format=fb.IN_S8, out_mode=fb.OUT_LOGMAG, log_mult=0.2 * np.log(10.0), power_gain=1.0)
    got = plan.exec_host(raw, frames).astype(np.float64)
    mag = oracle_rows(raw, n, n, frames, orc.IN_S8, orc.WIN_NONE, orc.OUT_MAG).astype(np.float64)
    ref = 10.0 * np.log(mag ** 2 + 1.0e-20) * 0.02
    ok = mag >= 1e-3 * np.sqrt(np.mean(mag ** 2, axis=1, keepdims=True))
    assert ok.mean() > 0.95
    assert np.max(np.abs(got - ref)[ok]) < 0.2 * np.log(10.0) / 10.0 * 2e-3     # the 2e-3 dB bound in the shader's units


@pytest.mark.parametrize("n", [16384, 65536])
def test_float_frames_above_8192_points(fbg, n):
    """FSEA_IN_F64C / F32C above 8192 points (the reference's float64 branch, src/nrf.c:607-609, has no such
    limit below its 131072-sample block): the plain fp64 path of fsea_bigf.cu, to 1e-11 of the oracle; and the
    nrf_fft operator at that size fed by a float64 buffer, as lua/fft-shifted.lua feeds it."""
    fb = fbg
    rng = np.random.default_rng(n)
    x = rng.normal(0, 1, 2 * 2 * n) + 0.5
    plan = fb.Plan(n, in_format=fb.IN_F64C, out_mode=fb.OUT_MAG, flags=fb.FLAG_DC_PATCH)
    got = plan.exec_host(x, 2)
    ref = np.stack([orc.engine_frame(x[f * 2 * n:], n, orc.IN_F64C, orc.WIN_NONE, orc.OUT_MAG, flags=orc.FLAG_DC_PATCH) for f in range(2)])
    assert got.dtype == np.float64 and np.max(np.abs(got - ref)) < 1e-11 * np.sqrt(np.mean(ref ** 2))
    assert np.array_equal(got[:, n // 2], got[:, n // 2 - 1])
    xf = x.astype(np.float32)
    plan = fb.Plan(n, in_format=fb.IN_F32C, window=fb.WIN_HANN, out_mode=fb.OUT_LOGMAG)
    got = plan.exec_host(xf, 2).astype(np.float64)
    ref = np.stack([orc.engine_frame(xf[f * 2 * n:], n, orc.IN_F32C, orc.WIN_HANN, orc.OUT_LOGMAG) for f in range(2)])
    assert np.max(np.abs(got - ref)) < 2e-3
    nrf = fb.nrf
    fft = nrf.fft_new(n, 2)
    buf = nrf.nut_buffer_new_f64(n, 2, x[:2 * n])
    nrf.fft_process(fft, buf)
    row = nrf.fft_get_buffer(fft).numpy()[:n]
    ref = orc.engine_frame(x, n, orc.IN_F64C, orc.WIN_NONE, orc.OUT_MAG, flags=orc.FLAG_DC_PATCH)
    assert np.max(np.abs(row - ref)) < 1e-11 * np.sqrt(np.mean(ref ** 2))
    nrf.fft_free(fft)


def test_fft_batch_all_blocks_of_the_big_capture(fbg):
    """BASELINE cfg1 (c/fft-batch.c over rfdata/rf-200.500-big.raw): ALL 100 blocks of the capture, where the file is
    at hand (the authoring container; its sha256 is checked) -- the reference's semantics (one 1024-point row
    per 262144-byte block, first 1024 samples, S = 10, no DC patch, c/fft-batch.c:54-102) and every one of the
    12800 non-overlapping frames -- against the oracle.  Elsewhere the committed fixture holds the first frame of
    every block (tests/golden/cfg1_block_heads.npz), so the 100 reference rows run on the GPU box too."""
    import hashlib
    fb = fbg
    n = 1024
    big = "/root/reference/rfdata/rf-200.500-big.raw"
    heads = np.load(os.path.join(os.path.dirname(__file__), "golden", "cfg1_block_heads.npz"))
    plan = fb.Plan(n, hop=n, in_format=fb.IN_S8, out_mode=fb.OUT_PIX, pix_scale=10.0)
    got = plan.exec_host(heads["heads"].reshape(-1), 100)                     # 100 blocks x first 1024 samples
    d = np.abs(got.astype(int) - heads["ref_pix"].astype(int))
    assert pix_close(d)
    assert got[0, 512] == 255 and int(heads["ref_pix"][0].sum()) == 23772       # SURVEY 8(c) goldens: block 0
    if not os.path.exists(big):
        return
    raw = np.fromfile(big, dtype=np.uint8)
    assert hashlib.sha256(raw.tobytes()).hexdigest().startswith("b7318ed7e926eacf") and raw.size == 100 * BLOCK
    assert np.array_equal(raw.reshape(100, BLOCK)[:, :2 * n], heads["heads"])
    frames = raw.size // (2 * n)
    assert frames == 12800
    mag = fb.Plan(n, hop=n, in_format=fb.IN_S8, out_mode=fb.OUT_MAG).exec_host(raw, frames)
    ref = orc.engine_batch(raw, n, n, frames, orc.IN_S8, orc.WIN_NONE, orc.OUT_MAG).astype(np.float64)
    assert rel_err(mag.astype(np.float64), ref) < MAG_TOL
    pix = plan.exec_host(raw, frames)
    refp = orc.engine_batch(raw, n, n, frames, orc.IN_S8, orc.WIN_NONE, orc.OUT_PIX, 10.0)
    d = np.abs(pix.astype(int) - refp.astype(int))
    assert pix_close(d)


def test_sweep_stitch(fbg, iq):
    """Sweep product: per-frequency pixel images (c/fft-batch-broad.c rows) max-blended side by side
    (c/fft-stitch-broad.c), against the oracle.  Four 'frequencies' = the four capture blocks."""
    fb = fbg
    rng = np.random.default_rng(4)
    imgs = rng.integers(0, 256, (7, 33, 256), dtype=np.uint8)
    for step in (256, 102, 40, 300):                        # WIDTH_STEP = N / (fs / step)
        assert np.array_equal(fb.stitch(imgs, step), orc.stitch(imgs, step))
    raw = iq["rf200big4"]
    n, rows = 256, 64
    plan = fb.Plan(n, in_format=fb.IN_S8, out_mode=fb.OUT_PIX, pix_scale=5.0, flags=fb.FLAG_DC_PATCH)
    images = np.stack([plan.exec_host(raw[b * BLOCK:], rows) for b in range(4)])
    ref_images = np.stack([oracle_rows(raw[b * BLOCK:], n, n, rows, orc.IN_S8, orc.WIN_NONE, orc.OUT_PIX, 5.0, orc.FLAG_DC_PATCH)
                           for b in range(4)])
    d = np.abs(fb.stitch(images, 102).astype(int) - orc.stitch(ref_images, 102).astype(int))
    assert pix_close(d)


def test_gate_mean(fbg, iq, golden):
    fb = fbg
    raw = iq["rf200big4"]
    plan = fb.Plan(256, hop=BLOCK // 2, in_format=fb.IN_S8, out_mode=fb.OUT_MAG)
    got = plan.gate_mean(raw, 4)
    rows = np.stack([orc.batch_row(raw[b * BLOCK:], 256, 5.0, True)[1] for b in range(4)])
    assert abs(got - orc.broad_gate_mean(rows)) < 1e-5 * got


@pytest.mark.parametrize("n", [16, 256, 1024, 4096])
def test_f64_and_f32c_inputs(fbg, n):
    fb = fbg
    rng = np.random.default_rng(n)
    frames = 3
    x = rng.standard_normal(2 * n * frames) * 0.3 + 0.5          # like the shifter's output (+0.5 offset)
    plan = fb.Plan(n, in_format=fb.IN_F64C, out_mode=fb.OUT_MAG, flags=fb.FLAG_DC_PATCH)
    got = plan.exec_host(x, frames)
    assert got.dtype == np.float64
    ref = np.stack([orc.nrf_fft_frame_f64(x[2 * n * f:], n) for f in range(frames)])
    assert rel_err(got, ref) < 1e-11
    xc = (x - 0.5).astype(np.float32)
    plan = fb.Plan(n, in_format=fb.IN_F32C, out_mode=fb.OUT_MAG)
    got = plan.exec_host(xc, frames)
    ref = oracle_rows(xc, n, n, frames, orc.IN_F32C, orc.WIN_NONE, orc.OUT_MAG)
    assert rel_err(got, ref) < MAG_TOL


def test_edge_cases(fbg):
    fb = fbg
    plan = fb.Plan(1024, in_format=fb.IN_U8)
    assert plan.exec_host(np.zeros(0, np.uint8), 0).shape == (0, 1024)      # empty batch
    one = plan.exec_host(synth_u8(1024), 1)                                   # single frame (half pair)
    assert one.shape == (1, 1024) and np.isfinite(one).all()
    with pytest.raises(ValueError):
        plan.exec_host(np.zeros(100, np.uint8), 1)                            # ragged: too short
    # extreme bytes: all 0x00 / 0xFF / 0x80
    for byte in (0, 255, 128):
        data = np.full(2048, byte, np.uint8)
        got = plan.exec_host(data, 1)
        ref = orc.engine_frame(data, 1024, orc.IN_U8, orc.WIN_NONE, orc.OUT_MAG)
        np.testing.assert_allclose(got[0], ref, atol=1e-3 * max(1.0, ref.max()) * 1e-2)


def test_large_batch_properties(fbg):
    """BASELINE-size batch through the device-resident call: Parseval per frame,
    frame independence (a frame's spectrum does not depend on its neighbours /
    position in the batch), determinism."""
    import torch
    fb = fbg
    n, frames = 1024, 1 << 16
    g = torch.Generator(device="cuda").manual_seed(1)
    d_in = torch.randint(0, 256, (frames * n * 2,), dtype=torch.uint8, device="cuda", generator=g)
    d_out = torch.empty((frames, n), dtype=torch.float32, device="cuda")
    plan = fb.Plan(n, in_format=fb.IN_U8, out_mode=fb.OUT_MAG)
    s = torch.cuda.current_stream().cuda_stream
    plan.exec_device(d_in.data_ptr(), d_out.data_ptr(), frames, s)
    torch.cuda.synchronize()
    x = d_in.view(frames, n, 2).to(torch.float64) / 256.0
    energy_t = (x ** 2).sum(dim=(1, 2)) * n
    energy_f = (d_out.to(torch.float64) ** 2).sum(dim=1)
    assert torch.max(torch.abs(energy_f / energy_t - 1)).item() < 1e-5        # Parseval
    # spot-check frames against the oracle, including the last (odd positions too)
    h_in = d_in.cpu().numpy()
    for f in (0, 1, 12345, frames - 2, frames - 1):
        ref = orc.engine_frame(h_in[f * 2 * n:], n, orc.IN_U8, orc.WIN_NONE, orc.OUT_MAG)
        k = np.arange(n) != n // 2
        assert rel_err(d_out[f].cpu().numpy()[k], ref[k]) < MAG_TOL
    # permuting frames permutes rows (independence) and re-running is bit-identical
    perm = torch.randperm(frames, device="cuda")
    d_in2 = d_in.view(frames, -1)[perm].contiguous()
    d_out2 = torch.empty_like(d_out)
    plan.exec_device(d_in2.data_ptr(), d_out2.data_ptr(), frames, s)
    torch.cuda.synchronize()
    assert torch.equal(d_out2, d_out[perm])


def test_nrf_fft_operator_matches_reference(fbg, iq, golden):
    """The reference's own call sequence (lua/fft.lua:42-50) against the history the
    compiled reference produced on the same four blocks."""
    fb = fbg
    nrf = fb.nrf
    raw = iq["rf200big4"]
    for n in (128, 1024):
        fft = nrf.fft_new(n, 6)
        for b in range(4):
            u8 = orc.device_unpack(raw[b * BLOCK:(b + 1) * BLOCK])
            buf = nrf.nut_buffer_new_u8(131072, 2, u8)
            nrf.fft_process(fft, buf)
        out = nrf.fft_get_buffer(fft)
        assert (out.type, out.length, out.channels) == (2, n * 6, 1)
        hist = out.numpy().reshape(6, n)
        ref = golden[f"ref_nrf_fft_hist_n{n}_h6"]
        assert rel_err(hist[:4], ref[:4]) < MAG_TOL
        assert not hist[4:].any()
        nrf.fft_shift(fft, 3.7)
        shifted = nrf.fft_get_buffer(fft).numpy().reshape(6, n)
        s = int(round(n / 3.7))
        assert np.array_equal(shifted[:, : n - s], hist[:, s:]) and not shifted[:, n - s:].any()
        nrf.fft_free(fft)


def test_nrf_fft_f64_chain(fbg, iq, golden):
    """lua/fft-shifted.lua:52-54: device -> freq shifter -> fft (float64 buffer, oversize)."""
    fb = fbg
    nrf = fb.nrf
    u8 = orc.device_unpack(iq["rf200big4"][:BLOCK])
    sh = nrf.freq_shifter_new(-10000, 5000000)
    buf = nrf.nut_buffer_new_u8(131072, 2, u8)
    nrf.freq_shifter_process(sh, buf)
    sb = nrf.freq_shifter_get_buffer(sh)
    fft = nrf.fft_new(1024, 2)
    nrf.fft_process(fft, sb)                       # 262144 "samples": clamped, not overflowed
    row = nrf.fft_get_buffer(fft).numpy().reshape(2, 1024)[0]
    assert rel_err(row, golden["ref_shifted_fft_n1024"]) < 1e-11
    nrf.fft_free(fft)
    nrf.freq_shifter_free(sh)


@pytest.mark.parametrize("taps,half_ampl", [(51, 200000), (97, 60000), (50, 100000)])
def test_iq_filter_matches_committed_golden(fbg, iq, golden, taps, half_ampl):
    """The same operator against vectors recorded from the compiled reference (tests/golden/make_golden.py):
    does not need oracle/_ref at run time."""
    from frequensea_b200 import nrf
    raw = orc.device_unpack(iq["rf200big4"])
    flt = nrf.iq_filter_new(5000000, half_ampl, taps)
    key = f"ref_iq_filter_t{taps}_f{half_ampl}"
    for b in range(3):
        buf = nrf.nut_buffer_new_u8(BLOCK // 2, 2, raw[b * BLOCK:(b + 1) * BLOCK])
        nrf.iq_filter_process(flt, buf)
        out = nrf.iq_filter_get_buffer(flt).numpy()
        assert np.array_equal(out[:4096], golden[key + "_head"][b])
        assert np.array_equal(out[-4096:], golden[key + "_tail"][b])
        assert out.sum() == golden[key + "_sum"][b]
    nrf.iq_filter_free(flt)


@pytest.mark.parametrize("taps,half_ampl", [(51, 200000), (97, 60000), (50, 100000)])
def test_iq_filter_bit_exact_and_chain(fbg, iq, taps, half_ampl):
    """nrf_iq_filter_* (src/nrf.c:644-776; lua/iq-tex-filtered.lua:43-46,
    lua/dvbt.lua:37) with the FIR evaluated by fsea_fir_iq_kernel -- bit-identical to the compiled
    reference on whole device blocks, state carried across blocks -- and chained into nrf_fft
    (float64 samples -> fp64 kernels) against the reference's own chain."""
    if not orc.have_ref():
        pytest.skip("oracle/_ref/libnrf_ref.so not built")
    from frequensea_b200 import nrf
    rate = 5000000
    raw = orc.device_unpack(iq["rf200big4"])                       # 4 device blocks, offset binary
    blocks = [raw[b * BLOCK:(b + 1) * BLOCK] for b in range(3)]
    ref = orc.ref_iq_filter_blocks(blocks, rate, half_ampl, taps)
    flt = nrf.iq_filter_new(rate, half_ampl, taps)
    fft = nrf.fft_new(1024, 4)
    for b, r in zip(blocks, ref):
        buf = nrf.nut_buffer_new_u8(BLOCK // 2, 2, b)
        nrf.iq_filter_process(flt, buf)
        out = nrf.iq_filter_get_buffer(flt)
        assert out.length == BLOCK // 2 and out.channels == 2 and out.type == nrf.NUT_BUFFER_F64
        assert np.array_equal(out.numpy(), r)
        nrf.fft_process(fft, out)
    hist = nrf.fft_get_buffer(fft).numpy().reshape(4, 1024)
    # the reference's chain on its own filtered blocks (newest row first)
    R = orc.ref()
    rfft = R.nrf_fft_new(1024, 4)
    for r in ref:
        nb = R.nut_buffer_new_f64(r.size // 2, 2, r.ctypes.data_as(__import__("ctypes").c_void_p))
        R.nrf_fft_process(rfft, nb)
        R.nut_buffer_free(nb)
    o = R.nrf_fft_get_buffer(rfft)
    rhist = orc.nut_to_numpy(o).reshape(4, 1024)
    R.nut_buffer_free(o)
    R.nrf_fft_free(rfft)
    assert rel_err(hist[:3], rhist[:3]) < 1e-11
    nrf.iq_filter_free(flt)
    nrf.fft_free(fft)


def test_capture_file_feed(fbg, iq, tmp_path):
    """a capture file as c/reader.c:25-49 writes it (raw int8 IQ, block after block)
    streamed through fsea_exec_file -- same rows as the host-buffer path and as the oracle; the
    reference replays such files block-wise (src/nrf.c:256-284)."""
    fb = fbg
    raw = iq["rf200big4"]                                  # 4 capture blocks
    path = str(tmp_path / "rf-200.500.raw")
    raw.tofile(path)
    n = 1024
    frames = raw.size // (2 * n)
    plan = fb.Plan(n, in_format=fb.IN_S8, out_mode=fb.OUT_MAG, flags=fb.FLAG_DC_PATCH)
    got = plan.exec_file(path, frames)
    assert np.array_equal(got, plan.exec_host(raw, frames))
    ref = oracle_rows(raw, n, n, 16, orc.IN_S8, orc.WIN_NONE, orc.OUT_MAG, flags=orc.FLAG_DC_PATCH)
    assert rel_err(got[:16], ref) < MAG_TOL
    # second block only, overlapping frames, pixels
    plan = fb.Plan(n, hop=512, in_format=fb.IN_S8, out_mode=fb.OUT_PIX, pix_scale=5.0, flags=fb.FLAG_DC_PATCH)
    got = plan.exec_file(path, 200, byte_offset=BLOCK)
    assert np.array_equal(got, plan.exec_host(raw[BLOCK:], 200))
    # many slots in flight: 65536-point frames from a larger file
    big = np.tile(raw, 8)
    bpath = str(tmp_path / "big.raw")
    big.tofile(bpath)
    plan = fb.Plan(65536, in_format=fb.IN_S8, out_mode=fb.OUT_LOGMAG)
    got = plan.exec_file(bpath, 64)
    assert np.array_equal(got, plan.exec_host(big, 64))
    with pytest.raises(RuntimeError):
        plan.exec_file(bpath, 65)                          # file too short


def test_sweep_engine_device_resident(fbg):
    """fsea_sweep_* (BASELINE cfg4 per GPU): raw IQ resident once, gate sum on the device
    (c/fft-batch-broad.c:80-98), pixel rows of the kept frequencies (:108-118), zeros for the others --
    against the oracle on the same bytes; host-fed and device-drawn (SURVEY 8(d) cfg4 source) frequencies,
    images in the engine's own store and placed in a stitched canvas (c/fft-stitch-broad.c, WIDTH_STEP = N)."""
    fb = fbg
    from frequensea_b200 import sweep
    from frequensea_b200.binding import Sweep
    n, rows, nf = 1024, 16, 6
    eng = Sweep(n, rows, nf, gate_rows=8, own_images=True)
    raws = []
    for f in range(nf):
        if f < 3:
            raw = sweep.synth_frequency(f, n, rows, quiet_every=3)
            eng.submit_host(f, raw)
        else:
            eng.submit_synth(f, f, 3)                       # drawn on the device: read the bytes back for the oracle
            raw = eng.read_raw(f)
        raws.append(raw)
    means, keep = eng.gate(nf)
    ref_means = np.array([orc.engine_batch(r, n, n, 8, orc.IN_S8, orc.WIN_NONE, orc.OUT_MAG).astype(np.float64).mean() for r in raws])
    np.testing.assert_allclose(means, ref_means, rtol=1e-5)
    assert keep.tolist() == [f % 3 != 2 for f in range(nf)]
    # the device-drawn source: quiet frequencies really are noise-only, the others carry their tone at bin (37 f + 11) mod N
    for f in (3, 4):
        mag = orc.engine_batch(raws[f], n, n, 1, orc.IN_S8, orc.WIN_NONE, orc.OUT_MAG, flags=orc.FLAG_DC_PATCH)[0]
        assert int(np.argmax(mag)) == (37 * f + 11) % n
    eng.render(nf)
    images = eng.read_images(0, nf)
    for f in range(nf):
        if not keep[f]:
            assert not images[f].any()
            continue
        ref = orc.engine_batch(raws[f], n, n, rows, orc.IN_S8, orc.WIN_NONE, orc.OUT_PIX, 5.0, flags=orc.FLAG_DC_PATCH)
        d = np.abs(images[f].astype(int) - ref.astype(int))
        assert pix_close(d)
    # the pixel pass in the reference's arithmetic: the oracle's pictures byte for byte (kept slots in runs: rows in groups)
    eng3 = Sweep(n, rows, nf, gate_rows=8, own_images=True, pixel_arith=fb.ARITH_F64)
    for f in range(nf):
        eng3.submit_host(f, raws[f])
    means3, keep3 = eng3.gate(nf)
    assert np.array_equal(keep3, keep)
    eng3.render(nf)
    images3 = eng3.read_images(0, nf)
    for f in range(nf):
        ref = orc.engine_batch(raws[f], n, n, rows, orc.IN_S8, orc.WIN_NONE, orc.OUT_PIX, 5.0, flags=orc.FLAG_DC_PATCH)
        assert np.array_equal(images3[f], ref if keep[f] else np.zeros_like(ref)), f
    # deterministic: a second engine fed the same bytes through submit_device / submit_host agrees bit for bit
    eng2 = Sweep(n, rows, nf, gate_rows=8, own_images=True)
    for f in range(nf):
        eng2.submit_host(f, raws[f])
    means2, keep2 = eng2.gate(nf)
    assert np.array_equal(means, means2) and np.array_equal(keep, keep2)
    eng2.render(nf)
    assert np.array_equal(eng2.read_images(0, nf), images)


def test_sweep_job_single_gpu(fbg):
    """The job (frequensea_b200/sweep.py) on one GPU: device engine == oracle stand-in; the canvas form
    (pixel rows stored straight into the stitched picture, what mode="p2p" does across GPUs) == gather + stitch.
    The N > 1 form runs over gloo in tests/test_shard_gloo.py and over NCCL in tools/sweep_demo.py / bench.py."""
    import torch
    from frequensea_b200 import sweep
    from conftest import OracleEngine
    n_freqs, n, rows, step = 6, 1024, 16, 400
    stream = lambda f: sweep.synth_frequency(f, n, rows, quiet_every=3)
    got = sweep.sweep(sweep.DeviceEngine(n, rows, n_freqs, gate_rows=8), n_freqs, n, rows, step, source=stream)
    ref = sweep.sweep(OracleEngine(n, rows, n_freqs, gate_rows=8), n_freqs, n, rows, step, source=stream)
    assert got["keep"].tolist() == ref["keep"].tolist() == [f % 3 != 2 for f in range(n_freqs)]
    np.testing.assert_allclose(got["means"], ref["means"], rtol=1e-5)
    gi, ri = got["images"].cpu().numpy(), ref["images"].numpy()
    d = np.abs(gi.astype(int) - ri.astype(int))
    assert pix_close(d)
    assert np.array_equal(got["stitched"].cpu().numpy(), orc.stitch(gi, step))     # stitch itself is bit-exact
    assert tuple(got["stitched"].shape) == (rows, n + (n_freqs - 1) * step)
    # WIDTH_STEP = N (c/fft-stitch-broad.c with SAMPLE_RATE = FREQUENCY_STEP): canvas stores == gather + stitch
    a = sweep.sweep(sweep.DeviceEngine(n, rows, n_freqs, gate_rows=8), n_freqs, n, rows, n, source=stream, mode="p2p")
    b = sweep.sweep(sweep.DeviceEngine(n, rows, n_freqs, gate_rows=8), n_freqs, n, rows, n, source=stream, mode="gather")
    assert torch.equal(a["stitched"], b["stitched"])


def test_mgpu_driver_one_process(fbg):
    """libfsea_mgpu.so (one process, ncclCommInitAll over the GPUs of the box): the cfg4 job on
    all visible GPUs (at most 2 here) == the same job on one GPU, bit for bit, in gather mode and with the
    pixel rows stored straight into device 0's canvas; and == the Python-driven engine.  Runs in a subprocess
    without torch (the driver links the system NCCL)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = min(2, fbg.device_count())
    for mode in ("gather", "p2p"):
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "sweep_mgpu.py"), "--gpus", str(g), "--freqs", "10", "--rows", "24",
                            "--fft-size", "4096", "--mode", mode, "--check", "--f32-gather-mib", "0"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        res = json.loads(r.stdout.strip().splitlines()[-1])
        assert res["kept"] == 9 and res["rows_rendered"] == 9 * 24          # frequency 9 is the quiet one (every 10th): gated off
        if g > 1:
            assert res["equals_one_gpu_job"]


def test_chain_fused_front_operators(fbg, iq):
    """fsea_chain_* (SURVEY 8(f) rank 3): device block -> frequency shifter / IQ low-pass -> nrf_fft row in one call on
    the device, only the N samples the transform reads going through the filter.  Against the SAME operators called one
    after the other through the reference's API (float64 nut_buffers of the whole block across the host, themselves
    bit-identical to the compiled reference: test_iq_filter_bit_exact_and_chain, test_freq_shifter_host_...): the
    filter's sums and the rotator's products are the same roundings, so the rows agree bit for bit, block after
    block (carried FIR tail, carried rotator phase)."""
    fb = fbg
    nrf = fb.nrf
    from frequensea_b200.binding import Chain
    n = 1024
    blocks = [orc.device_unpack(iq["rf200big4"][b * BLOCK:(b + 1) * BLOCK]) for b in range(4)]

    def unfused(make_front, process_front, get_front, free_front):
        fft = nrf.fft_new(n, 2)
        front = make_front()
        rows = []
        for blk in blocks:
            buf = nrf.nut_buffer_new_u8(131072, 2, blk)
            process_front(front, buf)
            nrf.fft_process(fft, get_front(front))
            rows.append(nrf.fft_get_buffer(fft).numpy()[:n].copy())
        free_front(front)
        nrf.fft_free(fft)
        return rows

    # IQ low-pass (lua/iq-tex-filtered.lua:43-46: 200 kHz, 51 taps; an even tap count as well)
    for taps, half in ((51, 200000), (50, 100000)):
        ref = unfused(lambda: nrf.iq_filter_new(5000000, half, taps), nrf.iq_filter_process, nrf.iq_filter_get_buffer, nrf.iq_filter_free)
        coef = nrf.fir_low_pass_coefficients(5000000, half, taps)[:taps]
        ch = Chain(n, fir_coefficients=coef)
        for blk, r in zip(blocks, ref):
            assert np.array_equal(ch.process(blk), r), taps
        assert ch.launches == 3 * len(blocks)                       # prepare + FIR + transform per block
    # frequency shifter (lua/fft-shifted.lua:42-54, -10 kHz at 5 MS/s)
    ref = unfused(lambda: nrf.freq_shifter_new(-10000, 5000000), nrf.freq_shifter_process, nrf.freq_shifter_get_buffer, nrf.freq_shifter_free)
    ch = Chain(n, shifter=(-10000, 5000000))
    for blk, r in zip(blocks, ref):
        assert np.array_equal(ch.process(blk), r)
    # both, against a float64 model of the intended composition (the reference's own shifter hands a buffer of twice
    # the block's length to whatever follows, src/nrf.c:849-852: scripts never chain the two)
    coef = nrf.fir_low_pass_coefficients(5000000, 200000, 51)[:51]
    ch = Chain(n, shifter=(-10000, 5000000), fir_coefficients=coef)
    step = 2 * np.pi * -10000 / 5000000.0
    c, s = 1.0, 0.0
    tail = np.zeros((50, 2))
    for blk in blocks[:2]:
        v = blk.astype(np.float64).reshape(-1, 2) / 256.0
        ph = np.empty((131072, 2))
        dc, ds = np.cos(step), np.sin(step)
        for j in range(131072):
            ph[j] = (c, s)
            c, s = c * dc - s * ds, c * ds + s * dc
        sh = np.stack([v[:, 0] * ph[:, 0] - v[:, 1] * ph[:, 1] + 0.5, v[:, 0] * ph[:, 1] + v[:, 1] * ph[:, 0] + 0.5], axis=1)
        line = np.concatenate([tail, sh[:n]])
        fi = np.array([np.dot(coef, line[i:i + 51, 0]) for i in range(n)])
        fq = np.array([np.dot(coef, line[i:i + 51, 1]) for i in range(n)])
        tail = sh[-50:]
        x = np.empty(2 * n)
        x[0::2], x[1::2] = fi, fq
        want = orc.engine_frame(x, n, orc.IN_F64C, orc.WIN_NONE, orc.OUT_MAG, flags=orc.FLAG_DC_PATCH)
        got = ch.process(blk)
        assert np.max(np.abs(got - want)) < 1e-11 * np.sqrt(np.mean(want ** 2))


def test_block_graph_filter_into_fft(fbg, iq):
    """nrf_block_connect / nrf_block_process (src/nrf.c:24-50): pushing a buffer into a filter block
    runs the filter, hands its result to the connected nrf_fft and frees it -- same history as calling
    the two operators by hand."""
    from frequensea_b200 import nrf
    raw = orc.device_unpack(iq["rf200big4"])
    blocks = [raw[b * BLOCK:(b + 1) * BLOCK] for b in range(2)]
    flt_a, fft_a = nrf.iq_filter_new(5000000, 150000, 33), nrf.fft_new(256, 4)
    flt_b, fft_b = nrf.iq_filter_new(5000000, 150000, 33), nrf.fft_new(256, 4)
    nrf.block_connect(flt_a, fft_a)                 # NRF_BLOCK is the first member of both objects
    for b in blocks:
        buf = nrf.nut_buffer_new_u8(BLOCK // 2, 2, b)
        nrf.block_process(flt_a, buf)
        nrf.iq_filter_process(flt_b, buf)
        nrf.fft_process(fft_b, nrf.iq_filter_get_buffer(flt_b))
    a, b_ = nrf.fft_get_buffer(fft_a).numpy(), nrf.fft_get_buffer(fft_b).numpy()
    assert a.any() and np.array_equal(a, b_)
    for o, free in ((flt_a, nrf.iq_filter_free), (flt_b, nrf.iq_filter_free), (fft_a, nrf.fft_free), (fft_b, nrf.fft_free)):
        free(o)
