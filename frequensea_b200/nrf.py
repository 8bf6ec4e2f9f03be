"""Python face of the reference's operator API (include/nrf.h, include/nut.h)
as exported by libfsea.so -- same names, argument meaning and ownership as the
reference's Lua globals (src/main.cpp:117-179,616-717,775-811), so tests read
like the reference's own scripts (lua/fft.lua:31-50):

    device = nrf.device_new(97, "rfdata/rf-200.500-big.raw")
    fft = nrf.fft_new(1024, 1024)
    samples = nrf.device_get_samples_buffer(device)
    nrf.fft_process(fft, samples)
    spectrum = nrf.fft_get_buffer(fft)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .binding import lib

NUT_BUFFER_U8, NUT_BUFFER_F64 = 1, 2          # src/nut.h:14-17
NRF_SAMPLES_LENGTH = 131072                   # src/nrf.h:20
NRF_BUFFER_SIZE_BYTES = 16 * 16384            # src/nrf.h:19


class NutBufferStruct(C.Structure):           # include/nut.h (src/nut.h:24-30)
    _fields_ = [("type", C.c_int), ("length", C.c_int), ("channels", C.c_int),
                ("size_bytes", C.c_int), ("data", C.c_void_p)]


_NB = C.POINTER(NutBufferStruct)
_ready = False


def _L():
    global _ready
    L = lib()
    if not _ready:
        vp, i, d = C.c_void_p, C.c_int, C.c_double
        L.nut_buffer_new_u8.argtypes = [i, i, vp]; L.nut_buffer_new_u8.restype = _NB
        L.nut_buffer_new_f64.argtypes = [i, i, vp]; L.nut_buffer_new_f64.restype = _NB
        L.nut_buffer_free.argtypes = [_NB]; L.nut_buffer_free.restype = None
        L.nut_buffer_copy.argtypes = [_NB]; L.nut_buffer_copy.restype = _NB
        L.nut_buffer_reduce.argtypes = [_NB, d]; L.nut_buffer_reduce.restype = _NB
        L.nut_buffer_clip.argtypes = [_NB, i, i]; L.nut_buffer_clip.restype = _NB
        L.nut_buffer_convert.argtypes = [_NB, i]; L.nut_buffer_convert.restype = _NB
        L.nut_buffer_append.argtypes = [_NB, _NB]; L.nut_buffer_append.restype = None
        L.nut_buffer_set_data.argtypes = [_NB, _NB]; L.nut_buffer_set_data.restype = None
        L.nut_buffer_get_u8.argtypes = [_NB, i]; L.nut_buffer_get_u8.restype = C.c_uint8
        L.nut_buffer_get_f64.argtypes = [_NB, i]; L.nut_buffer_get_f64.restype = d
        L.nut_buffer_set_u8.argtypes = [_NB, i, C.c_uint8]; L.nut_buffer_set_u8.restype = None
        L.nut_buffer_set_f64.argtypes = [_NB, i, d]; L.nut_buffer_set_f64.restype = None
        L.nut_buffer_save.argtypes = [_NB, C.c_char_p]; L.nut_buffer_save.restype = None
        L.nrf_block_connect.argtypes = [vp, vp]; L.nrf_block_connect.restype = None
        L.nrf_block_process.argtypes = [vp, _NB]; L.nrf_block_process.restype = None
        L.nrf_device_new.argtypes = [d, C.c_char_p]; L.nrf_device_new.restype = vp
        L.nrf_device_set_frequency.argtypes = [vp, d]; L.nrf_device_set_frequency.restype = d
        L.nrf_device_set_paused.argtypes = [vp, i]; L.nrf_device_set_paused.restype = None
        L.nrf_device_step.argtypes = [vp]; L.nrf_device_step.restype = None
        L.nrf_device_get_samples_buffer.argtypes = [vp]; L.nrf_device_get_samples_buffer.restype = _NB
        L.nrf_device_free.argtypes = [vp]; L.nrf_device_free.restype = None
        L.nrf_fft_new.argtypes = [i, i]; L.nrf_fft_new.restype = vp
        L.nrf_fft_shift.argtypes = [vp, d]; L.nrf_fft_shift.restype = None
        L.nrf_fft_process.argtypes = [vp, _NB]; L.nrf_fft_process.restype = None
        L.nrf_fft_get_buffer.argtypes = [vp]; L.nrf_fft_get_buffer.restype = _NB
        L.nrf_fft_free.argtypes = [vp]; L.nrf_fft_free.restype = None
        L.nrf_freq_shifter_new.argtypes = [i, i]; L.nrf_freq_shifter_new.restype = vp
        L.nrf_freq_shifter_process.argtypes = [vp, _NB]; L.nrf_freq_shifter_process.restype = None
        L.nrf_freq_shifter_get_buffer.argtypes = [vp]; L.nrf_freq_shifter_get_buffer.restype = _NB
        L.nrf_freq_shifter_free.argtypes = [vp]; L.nrf_freq_shifter_free.restype = None
        L.nrf_iq_filter_new.argtypes = [i, i, i]; L.nrf_iq_filter_new.restype = vp
        L.nrf_iq_filter_process.argtypes = [vp, _NB]; L.nrf_iq_filter_process.restype = None
        L.nrf_iq_filter_get_buffer.argtypes = [vp]; L.nrf_iq_filter_get_buffer.restype = _NB
        L.nrf_iq_filter_free.argtypes = [vp]; L.nrf_iq_filter_free.restype = None
        L.nrf_fir_get_low_pass_coefficients.argtypes = [i, i, i]; L.nrf_fir_get_low_pass_coefficients.restype = C.POINTER(d)
        _ready = True
    return L


class NutBuffer:
    """Owning handle of a nut_buffer* (the Lua binding frees through __gc,
    src/main.cpp:175-179,1095)."""

    def __init__(self, ptr, own: bool = True):
        self.ptr, self._own = ptr, own

    @property
    def type(self): return self.ptr.contents.type
    @property
    def length(self): return self.ptr.contents.length
    @property
    def channels(self): return self.ptr.contents.channels
    @property
    def size_bytes(self): return self.ptr.contents.size_bytes

    def numpy(self) -> np.ndarray:
        b = self.ptr.contents
        count = b.length * b.channels
        ct = C.c_uint8 if b.type == NUT_BUFFER_U8 else C.c_double
        if count == 0:
            return np.empty(0, dtype=np.uint8 if b.type == NUT_BUFFER_U8 else np.float64)
        return np.ctypeslib.as_array(C.cast(b.data, C.POINTER(ct)), shape=(count,)).copy()

    def free(self):
        if self._own and self.ptr:
            _L().nut_buffer_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def nut_buffer_new_u8(length, channels, data=None) -> NutBuffer:
    p = None
    if data is not None:
        data = np.ascontiguousarray(data, dtype=np.uint8)
        p = data.ctypes.data_as(C.c_void_p)
    return NutBuffer(_L().nut_buffer_new_u8(length, channels, p))


def nut_buffer_new_f64(length, channels, data=None) -> NutBuffer:
    p = None
    if data is not None:
        data = np.ascontiguousarray(data, dtype=np.float64)
        p = data.ctypes.data_as(C.c_void_p)
    return NutBuffer(_L().nut_buffer_new_f64(length, channels, p))


def nut_buffer_copy(b): return NutBuffer(_L().nut_buffer_copy(b.ptr))
def nut_buffer_reduce(b, pct): return NutBuffer(_L().nut_buffer_reduce(b.ptr, pct))
def nut_buffer_clip(b, off, length): return NutBuffer(_L().nut_buffer_clip(b.ptr, off, length))
def nut_buffer_convert(b, new_type): return NutBuffer(_L().nut_buffer_convert(b.ptr, new_type))
def nut_buffer_append(dst, src): _L().nut_buffer_append(dst.ptr, src.ptr)
def nut_buffer_save(b, fname): _L().nut_buffer_save(b.ptr, fname.encode())


def device_new(freq_mhz: float, data_file: str):
    return _L().nrf_device_new(freq_mhz, data_file.encode() if data_file else None)


def device_set_frequency(dev, f): return _L().nrf_device_set_frequency(dev, f)
def device_set_paused(dev, paused): _L().nrf_device_set_paused(dev, int(bool(paused)))
def device_step(dev): _L().nrf_device_step(dev)
def device_get_samples_buffer(dev): return NutBuffer(_L().nrf_device_get_samples_buffer(dev))
def device_free(dev): _L().nrf_device_free(dev)
def block_connect(a, b): _L().nrf_block_connect(a, b)
def block_process(block, b: NutBuffer): _L().nrf_block_process(block, b.ptr)

def fft_new(fft_size: int, fft_history_size: int): return _L().nrf_fft_new(fft_size, fft_history_size)
def fft_shift(fft, d: float): _L().nrf_fft_shift(fft, d)
def fft_process(fft, buffer: NutBuffer): _L().nrf_fft_process(fft, buffer.ptr)
def fft_get_buffer(fft) -> NutBuffer: return NutBuffer(_L().nrf_fft_get_buffer(fft))
def fft_free(fft): _L().nrf_fft_free(fft)

def freq_shifter_new(off: int, rate: int): return _L().nrf_freq_shifter_new(off, rate)
def freq_shifter_process(sh, b: NutBuffer): _L().nrf_freq_shifter_process(sh, b.ptr)
def freq_shifter_get_buffer(sh) -> NutBuffer: return NutBuffer(_L().nrf_freq_shifter_get_buffer(sh))
def freq_shifter_free(sh): _L().nrf_freq_shifter_free(sh)

# lua/iq-tex-filtered.lua:43-46: filter = nrf_iq_filter_new(sample_rate, 200e3, 51); nrf_iq_filter_process(filter, samples)
def iq_filter_new(sample_rate: int, half_ampl_freq: int, kernel_length: int): return _L().nrf_iq_filter_new(sample_rate, half_ampl_freq, kernel_length)
def iq_filter_process(f, b: NutBuffer): _L().nrf_iq_filter_process(f, b.ptr)
def iq_filter_get_buffer(f) -> NutBuffer: return NutBuffer(_L().nrf_iq_filter_get_buffer(f))
def iq_filter_free(f): _L().nrf_iq_filter_free(f)


def fir_low_pass_coefficients(sample_rate: int, half_ampl_freq: int, length: int) -> np.ndarray:
    n = length + (length + 1) % 2
    p = _L().nrf_fir_get_low_pass_coefficients(sample_rate, half_ampl_freq, length)
    out = np.ctypeslib.as_array(p, shape=(n,)).copy()
    C.CDLL(None).free(p)
    return out
