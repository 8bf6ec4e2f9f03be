"""ctypes front-end for the CPU oracle (oracle/liboracle.so) and, when present,
for the compiled reference (oracle/_ref/libnrf_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(frequensea_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "libnrf_ref.so")

IN_U8, IN_S8, IN_F32C, IN_F64C = 0, 1, 2, 3
WIN_NONE, WIN_HANN = 0, 1
OUT_MAG, OUT_LOGMAG, OUT_PIX = 0, 1, 2
FLAG_DC_PATCH = 1

NRF_BUFFER_SIZE_BYTES = 16 * 16384   # src/nrf.h:19
NRF_SAMPLES_LENGTH = 131072          # src/nrf.h:20


def build(force: bool = False) -> None:
    """Compile liboracle.so (and _ref/libnrf_ref.so where /root/reference exists)."""
    if force or not os.path.exists(_LIB) or \
            os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "oracle.c")):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/src/nrf.c") and (force or not os.path.exists(_REF)):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, stdout=subprocess.DEVNULL)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        vp, i, d, l = C.c_void_p, C.c_int, C.c_double, C.c_long
        L.orc_device_unpack.argtypes = [vp, vp, C.c_size_t, i]
        L.orc_dft_forward.argtypes = [i, vp, vp]
        L.orc_dft_naive.argtypes = [i, vp, vp]
        L.orc_nrf_fft_frame_u8.argtypes = [i, vp, vp]
        L.orc_nrf_fft_frame_f64.argtypes = [i, vp, vp]
        L.orc_history_push.argtypes = [vp, i, i, vp]
        L.orc_fft_shift.argtypes = [vp, i, i, d]
        L.orc_batch_row.argtypes = [i, vp, d, i, vp, vp]
        L.orc_broad_gate_mean.argtypes = [vp, i, i]
        L.orc_broad_gate_mean.restype = d
        L.orc_scope_row.argtypes = [i, vp, d, vp]
        L.orc_stitch.argtypes = [vp, i, i, i, i, vp]
        L.orc_stitch.restype = None
        L.orc_engine_frame.argtypes = [i, i, i, i, d, i, vp, vp]
        L.orc_engine_batch.argtypes = [i, i, i, i, i, d, i, vp, l, vp, i]
        for f in ("orc_device_unpack", "orc_dft_forward", "orc_dft_naive", "orc_nrf_fft_frame_u8",
                  "orc_nrf_fft_frame_f64", "orc_history_push", "orc_fft_shift", "orc_batch_row",
                  "orc_scope_row", "orc_engine_frame", "orc_engine_batch"):
            getattr(L, f).restype = None
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def device_unpack(raw: np.ndarray, convert: bool = True) -> np.ndarray:
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    out = np.empty_like(raw)
    lib().orc_device_unpack(_p(raw), _p(out), raw.size, int(convert))
    return out


def dft_forward(x: np.ndarray, naive: bool = False) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.complex128)
    out = np.empty_like(x)
    (lib().orc_dft_naive if naive else lib().orc_dft_forward)(x.size, _p(x), _p(out))
    return out


def nrf_fft_frame_u8(samples: np.ndarray, n: int) -> np.ndarray:
    samples = np.ascontiguousarray(samples[:2 * n], dtype=np.uint8)
    mag = np.empty(n, dtype=np.float64)
    lib().orc_nrf_fft_frame_u8(n, _p(samples), _p(mag))
    return mag


def nrf_fft_frame_f64(samples: np.ndarray, n: int) -> np.ndarray:
    samples = np.ascontiguousarray(samples[:2 * n], dtype=np.float64)
    mag = np.empty(n, dtype=np.float64)
    lib().orc_nrf_fft_frame_f64(n, _p(samples), _p(mag))
    return mag


def history_push(history: np.ndarray, row0: np.ndarray) -> None:
    rows, n = history.shape
    assert history.flags.c_contiguous and history.dtype == np.float64
    row0 = np.ascontiguousarray(row0, dtype=np.float64)
    lib().orc_history_push(_p(history), n, rows, _p(row0))


def fft_shift(history: np.ndarray, d: float) -> None:
    rows, n = history.shape
    assert history.flags.c_contiguous and history.dtype == np.float64
    lib().orc_fft_shift(_p(history), n, rows, float(d))


def batch_row(raw: np.ndarray, n: int, scale: float, dc_patch: bool):
    raw = np.ascontiguousarray(raw[:2 * n], dtype=np.uint8)
    pix = np.empty(n, dtype=np.uint8)
    cplx = np.empty(n, dtype=np.complex128)
    lib().orc_batch_row(n, _p(raw), float(scale), int(dc_patch), _p(pix), _p(cplx))
    return pix, cplx


def broad_gate_mean(cplx_rows: np.ndarray) -> float:
    cplx_rows = np.ascontiguousarray(cplx_rows, dtype=np.complex128)
    rows, n = cplx_rows.shape
    return float(lib().orc_broad_gate_mean(_p(cplx_rows), n, rows))


def stitch(images: np.ndarray, width_step: int) -> np.ndarray:
    images = np.ascontiguousarray(images, dtype=np.uint8)
    n_images, rows, width = images.shape
    out = np.empty((rows, width + (n_images - 1) * width_step), dtype=np.uint8)
    lib().orc_stitch(_p(images), n_images, rows, width, width_step, _p(out))
    return out


def scope_row(raw: np.ndarray, n: int, intensity: float) -> np.ndarray:
    raw = np.ascontiguousarray(raw[:2 * n], dtype=np.uint8)
    out = np.empty(n, dtype=np.float64)
    lib().orc_scope_row(n, _p(raw), float(intensity), _p(out))
    return out


_IN_DT = {IN_U8: (np.uint8, 2), IN_S8: (np.uint8, 2), IN_F32C: (np.float32, 2), IN_F64C: (np.float64, 2)}


def engine_frame(frame: np.ndarray, n: int, in_format=IN_U8, window=WIN_NONE, out_mode=OUT_MAG,
                 pix_scale=10.0, flags=0) -> np.ndarray:
    dt, per = _IN_DT[in_format]
    frame = np.ascontiguousarray(frame, dtype=dt).ravel()[:per * n]
    out = np.empty(n, dtype=np.uint8 if out_mode == OUT_PIX else np.float64)
    lib().orc_engine_frame(n, in_format, window, out_mode, float(pix_scale), flags, _p(frame), _p(out))
    return out


def engine_batch(data: np.ndarray, n: int, hop: int, n_frames: int, in_format=IN_U8,
                 window=WIN_NONE, out_mode=OUT_MAG, pix_scale=10.0, flags=0, threads=1) -> np.ndarray:
    dt, per = _IN_DT[in_format]
    data = np.ascontiguousarray(data, dtype=dt).ravel()
    need = ((n_frames - 1) * hop + n) * per if n_frames > 0 else 0
    assert data.size >= need, (data.size, need)
    out = np.empty((n_frames, n), dtype=np.uint8 if out_mode == OUT_PIX else np.float32)
    if n_frames:
        lib().orc_engine_batch(n, hop, in_format, window, out_mode, float(pix_scale), flags,
                               _p(data), n_frames, _p(out), threads)
    return out


# --------------------------------------------------------------------------
# The compiled reference (src/nrf.c + src/nut.c, third-party libs stubbed).
# --------------------------------------------------------------------------
class NutBuffer(C.Structure):            # src/nut.h:24-30
    _fields_ = [("type", C.c_int), ("length", C.c_int), ("channels", C.c_int),
                ("size_bytes", C.c_int), ("data", C.c_void_p)]


def have_ref() -> bool:
    if not os.path.exists(_REF) and os.path.exists("/root/reference/src/nrf.c"):
        build()
    return os.path.exists(_REF)


_ref = None


def ref() -> C.CDLL:
    """The reference's own nrf_fft_* / nut_buffer_* entry points."""
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libnrf_ref.so not built (needs /root/reference)")
        R = C.CDLL(_REF)
        vp, i, d = C.c_void_p, C.c_int, C.c_double
        NB = C.POINTER(NutBuffer)
        R.nut_buffer_new_u8.argtypes = [i, i, vp]; R.nut_buffer_new_u8.restype = NB
        R.nut_buffer_new_f64.argtypes = [i, i, vp]; R.nut_buffer_new_f64.restype = NB
        R.nut_buffer_free.argtypes = [NB]; R.nut_buffer_free.restype = None
        R.nut_buffer_convert.argtypes = [NB, i]; R.nut_buffer_convert.restype = NB
        R.nut_buffer_clip.argtypes = [NB, i, i]; R.nut_buffer_clip.restype = NB
        R.nut_buffer_reduce.argtypes = [NB, d]; R.nut_buffer_reduce.restype = NB
        R.nut_buffer_append.argtypes = [NB, NB]; R.nut_buffer_append.restype = None
        R.nrf_fft_new.argtypes = [i, i]; R.nrf_fft_new.restype = vp
        R.nrf_fft_process.argtypes = [vp, NB]; R.nrf_fft_process.restype = None
        R.nrf_fft_shift.argtypes = [vp, d]; R.nrf_fft_shift.restype = None
        R.nrf_fft_get_buffer.argtypes = [vp]; R.nrf_fft_get_buffer.restype = NB
        R.nrf_fft_free.argtypes = [vp]; R.nrf_fft_free.restype = None
        R.nrf_device_new.argtypes = [d, C.c_char_p]; R.nrf_device_new.restype = vp
        R.nrf_device_set_paused.argtypes = [vp, i]; R.nrf_device_set_paused.restype = None
        R.nrf_device_step.argtypes = [vp]; R.nrf_device_step.restype = None
        R.nrf_device_get_samples_buffer.argtypes = [vp]; R.nrf_device_get_samples_buffer.restype = NB
        R.nrf_device_free.argtypes = [vp]; R.nrf_device_free.restype = None
        R.nrf_freq_shifter_new.argtypes = [i, i]; R.nrf_freq_shifter_new.restype = vp
        R.nrf_freq_shifter_process.argtypes = [vp, NB]; R.nrf_freq_shifter_process.restype = None
        R.nrf_freq_shifter_get_buffer.argtypes = [vp]; R.nrf_freq_shifter_get_buffer.restype = NB
        R.nrf_iq_filter_new.argtypes = [i, i, i]; R.nrf_iq_filter_new.restype = vp
        R.nrf_iq_filter_process.argtypes = [vp, NB]; R.nrf_iq_filter_process.restype = None
        R.nrf_iq_filter_get_buffer.argtypes = [vp]; R.nrf_iq_filter_get_buffer.restype = NB
        R.nrf_iq_filter_free.argtypes = [vp]; R.nrf_iq_filter_free.restype = None
        R.nrf_fir_get_low_pass_coefficients.argtypes = [i, i, i]; R.nrf_fir_get_low_pass_coefficients.restype = C.POINTER(d)
        _ref = R
    return _ref


def nut_to_numpy(nb) -> np.ndarray:
    """Copy a nut_buffer* (either library -- same struct) into a numpy array."""
    b = nb.contents
    count = b.length * b.channels
    if b.type == 1:
        return np.ctypeslib.as_array(C.cast(b.data, C.POINTER(C.c_uint8)), shape=(count,)).copy()
    return np.ctypeslib.as_array(C.cast(b.data, C.POINTER(C.c_double)), shape=(count,)).copy()


def ref_nrf_fft_rows(samples_u8: np.ndarray, n: int, history: int, frames) -> np.ndarray:
    """Run the REFERENCE nrf_fft over the given frames (each a u8 offset-binary
    array of >= 2n bytes); returns the final history buffer (history x n)."""
    R = ref()
    fft = R.nrf_fft_new(n, history)
    for fr in frames:
        fr = np.ascontiguousarray(fr, dtype=np.uint8)
        nb = R.nut_buffer_new_u8(fr.size // 2, 2, _p(fr))
        R.nrf_fft_process(fft, nb)
        R.nut_buffer_free(nb)
    out = R.nrf_fft_get_buffer(fft)
    arr = nut_to_numpy(out).reshape(history, n)
    R.nut_buffer_free(out)
    R.nrf_fft_free(fft)
    return arr


def ref_iq_filter_blocks(blocks, sample_rate: int, half_ampl_freq: int, kernel_length: int, f64: bool = False):
    """Run the REFERENCE nrf_iq_filter (src/nrf.c:736-776) over consecutive sample blocks (u8
    offset-binary interleaved IQ, or float64 when f64); returns the get_buffer() result per block."""
    R = ref()
    flt = R.nrf_iq_filter_new(sample_rate, half_ampl_freq, kernel_length)
    outs = []
    for b in blocks:
        if f64:
            b = np.ascontiguousarray(b, dtype=np.float64)
            nb = R.nut_buffer_new_f64(b.size // 2, 2, _p(b))
        else:
            b = np.ascontiguousarray(b, dtype=np.uint8)
            nb = R.nut_buffer_new_u8(b.size // 2, 2, _p(b))
        R.nrf_iq_filter_process(flt, nb)
        R.nut_buffer_free(nb)
        o = R.nrf_iq_filter_get_buffer(flt)
        outs.append(nut_to_numpy(o))
        R.nut_buffer_free(o)
    R.nrf_iq_filter_free(flt)
    return outs


def ref_fir_coefficients(sample_rate: int, half_ampl_freq: int, length: int) -> np.ndarray:
    n = length + (length + 1) % 2
    p = ref().nrf_fir_get_low_pass_coefficients(sample_rate, half_ampl_freq, length)
    out = np.ctypeslib.as_array(p, shape=(n,)).copy()
    C.CDLL(None).free(p)
    return out
