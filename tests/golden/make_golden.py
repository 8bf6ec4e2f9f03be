#!/usr/bin/env python3
"""Generate tests/golden/*.npz -- run in the authoring container only
(needs /root/reference for the rfdata captures and, for the `ref_*` arrays,
oracle/_ref/libnrf_ref.so = the reference's own src/nrf.c + src/nut.c).

    python tests/golden/make_golden.py

Contents
  iq_fixtures.npz   raw capture bytes (two's-complement int8 IQ, as written by
                    c/reader.c:37-46): rf-433.000.raw (1 block) and the first
                    4 blocks of rf-200.500-big.raw
  golden_ref.npz    float64 results on those bytes, computed two independent ways:
                    np_*  : numpy (pocketfft) restatement of SURVEY.md 8(a)
                    ref_* : the compiled reference itself (nrf_fft_new/process/
                            get_buffer, nrf_fft_shift, nrf_freq_shifter_*,
                            nut_buffer_*, nrf_iq_filter_*), FFTW role stubbed by oracle/stubs
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
REF = "/root/reference/rfdata"
BLOCK = 262144


def np_unpack(raw):                       # src/nrf.c:102-105
    return ((raw.astype(np.int32) + 128) % 256).astype(np.uint8)


def np_spectrum(u8, n):                   # src/nrf.c:599-615 (before magnitude)
    x = (u8[0:2 * n:2].astype(np.float64) + 1j * u8[1:2 * n:2].astype(np.float64)) / 256.0
    x = x * (1 - 2 * (np.arange(n) % 2))
    return np.fft.fft(x)


def np_mag(u8, n):                        # src/nrf.c:619-630
    m = np.abs(np_spectrum(u8, n))
    m[n // 2] = m[n // 2 - 1]
    return m


def np_pix(u8, n, scale, patch):          # c/fft-batch.c:85-92, c/fft-batch-broad.c:108-118
    X = np_spectrum(u8, n)
    db = 10.0 * np.log10(X.real ** 2 + X.imag ** 2 + 1e-20) * scale
    p = np.clip(np.trunc(db), 0, 255).astype(np.uint8)
    if patch:
        p[n // 2] = p[n // 2 - 1]
    return p


def main():
    r433 = np.fromfile(os.path.join(REF, "rf-433.000.raw"), dtype=np.uint8)
    big = np.fromfile(os.path.join(REF, "rf-200.500-big.raw"), dtype=np.uint8)[:4 * BLOCK]
    np.savez_compressed(os.path.join(HERE, "iq_fixtures.npz"), rf433=r433, rf200big4=big)

    out = {}
    u433, ubig = np_unpack(r433), np_unpack(big)
    for name, u in (("rf433", u433), ("rf200big", ubig)):
        for n in (128, 1024):
            frames = u[:BLOCK].reshape(-1, 2 * n)
            mags = np.stack([np_mag(f, n) for f in frames])
            out[f"np_mag_{name}_n{n}_first8"] = mags[:8]
            out[f"np_magsum_{name}_n{n}"] = mags.sum(axis=1)          # every frame of block 0
            out[f"np_dcbin_{name}_n{n}"] = np.abs(np_spectrum(frames[0], n))[n // 2]
    out["np_pix_batch_n1024"] = np.stack([np_pix(ubig[b * BLOCK:], 1024, 10.0, False) for b in range(4)])
    out["np_pix_broad_n256"] = np.stack([np_pix(ubig[b * BLOCK:], 256, 5.0, True) for b in range(4)])
    full = np_unpack(np.fromfile(os.path.join(REF, "rf-200.500-big.raw"), dtype=np.uint8))
    rows = np.stack([np_spectrum(full[b * BLOCK:], 256) for b in range(100)])
    out["np_gate_mean_n256_100rows"] = np.array(np.abs(rows).mean())   # c/fft-batch-broad.c:80-98

    from oracle import orc
    if orc.have_ref():
        R = orc.ref()
        # nrf_fft over the 4 blocks, exactly as lua/fft.lua drives it (1 FFT per block)
        for n in (128, 1024):
            blocks = [ubig[b * BLOCK:(b + 1) * BLOCK] for b in range(4)]
            out[f"ref_nrf_fft_hist_n{n}_h6"] = orc.ref_nrf_fft_rows(None, n, 6, blocks)
        # nrf_fft_shift on a known ramp image
        img = (np.arange(8 * 128, dtype=np.float64).reshape(8, 128) + 1.0)
        for d in (3.7, -9.0, 0.9, 1e9):
            fft = R.nrf_fft_new(128, 8)
            import ctypes as C
            class F(C.Structure):
                _fields_ = [("block", C.c_byte * 112), ("fft_size", C.c_int), ("hist", C.c_int), ("buffer", C.POINTER(C.c_double))]
            f = C.cast(fft, C.POINTER(F)).contents
            assert f.fft_size == 128 and f.hist == 8
            C.memmove(f.buffer, img.ctypes.data, img.nbytes)
            R.nrf_fft_shift(fft, d)
            nb = R.nrf_fft_get_buffer(fft)
            out[f"ref_shift_d{d}"] = orc.nut_to_numpy(nb).reshape(8, 128)
            R.nut_buffer_free(nb)
            R.nrf_fft_free(fft)
        out["ref_shift_input"] = img
        # freq shifter -> fft (lua/fft-shifted.lua:52-54), block 0, N = 1024
        sh = R.nrf_freq_shifter_new(-10000, 5000000)
        nb = R.nut_buffer_new_u8(131072, 2, ubig[:BLOCK].ctypes.data)
        R.nrf_freq_shifter_process(sh, nb)
        sb = R.nrf_freq_shifter_get_buffer(sh)
        arr = orc.nut_to_numpy(sb)
        out["ref_shifter_out_first4096"] = arr[:4096]
        out["ref_shifter_len_channels"] = np.array([sb.contents.length, sb.contents.channels])
        # NB: the reference's own nrf_fft_process would overflow fft_in on this
        # oversize buffer; feed it the first 131072 samples only.
        clip = R.nut_buffer_new_f64(131072, 2, arr.ctypes.data)
        fft = R.nrf_fft_new(1024, 2)
        R.nrf_fft_process(fft, clip)
        hb = R.nrf_fft_get_buffer(fft)
        out["ref_shifted_fft_n1024"] = orc.nut_to_numpy(hb).reshape(2, 1024)[0]
        for b in (nb, sb, clip, hb):
            R.nut_buffer_free(b)
        R.nrf_fft_free(fft)
        # nrf_iq_filter over the first 3 device blocks (lua/iq-tex-filtered.lua:43-46: 200 kHz, 51 taps; lua/dvbt.lua:37:
        # 60 kHz, 97 taps; an even tap count, where the reference keeps length + 1 coefficients but uses length)
        blocks = [ubig[b * BLOCK:(b + 1) * BLOCK] for b in range(3)]
        for taps, half in ((51, 200000), (97, 60000), (50, 100000)):
            outs = orc.ref_iq_filter_blocks(blocks, 5000000, half, taps)
            out[f"ref_fir_coef_t{taps}_f{half}"] = orc.ref_fir_coefficients(5000000, half, taps)
            out[f"ref_iq_filter_t{taps}_f{half}_head"] = np.stack([o[:4096] for o in outs])         # first 2048 samples of each block
            out[f"ref_iq_filter_t{taps}_f{half}_tail"] = np.stack([o[-4096:] for o in outs])
            out[f"ref_iq_filter_t{taps}_f{half}_sum"] = np.array([o.sum() for o in outs])
    np.savez_compressed(os.path.join(HERE, "golden_ref.npz"), **out)
    # BASELINE cfg1 (c/fft-batch.c over rf-200.500-big.raw): the reference transforms the first 1024 samples of every
    # 262144-byte block -- those 2048 bytes of each of the 100 blocks, and the pixel rows c/fft-batch.c:61-92 makes of
    # them (oracle restatement: S = 10, no DC patch), so that the 100 rows run where the 26 MB capture is not at hand
    full = np.fromfile(os.path.join(REF, "rf-200.500-big.raw"), dtype=np.uint8)
    assert hashlib.sha256(full.tobytes()).hexdigest().startswith("b7318ed7e926eacf")
    heads = np.ascontiguousarray(full.reshape(100, BLOCK)[:, :2048])
    ref_pix = np.stack([orc.engine_frame(h, 1024, orc.IN_S8, orc.WIN_NONE, orc.OUT_PIX, 10.0, 0) for h in heads]).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "cfg1_block_heads.npz"), heads=heads, ref_pix=ref_pix)
    print("cfg1_block_heads", heads.shape, ref_pix.shape, int(ref_pix[0].sum()), int(ref_pix[0, 512]))
    for k, v in out.items():
        print(k, getattr(v, "shape", None))


if __name__ == "__main__":
    main()
