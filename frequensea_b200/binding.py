"""ctypes binding of include/fsea.h (libfsea.so)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfsea.so")


def _use_library(path: str) -> None:
    """Test / tooling hook (tests/emu, tools/sweep.py A/B builds): bind this module to another build of
    the library BEFORE the first call.  The product never calls it: frequensea_b200 resolves its own
    libfsea.so and nothing else -- no environment variable can point it at a CPU build."""
    global _LIB_PATH, _lib
    _LIB_PATH = path
    _lib = None

IN_U8, IN_S8, IN_F32C, IN_F64C = 0, 1, 2, 3
WIN_NONE, WIN_HANN = 0, 1
OUT_MAG, OUT_LOGMAG, OUT_PIX = 0, 1, 2
FLAG_DC_PATCH = 1
ARITH_AUTO, ARITH_F32, ARITH_F32X2, ARITH_F32X2_FRAME, ARITH_F32X2_RADIX64, ARITH_F64 = 0, 1, 2, 3, 4, 5

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_CUDA, ERR_NOMEM = 0, -1, -2, -3, -4, -5


class FseaConfig(C.Structure):
    _fields_ = [("fft_size", C.c_int), ("hop", C.c_int), ("in_format", C.c_int),
                ("window", C.c_int), ("out_mode", C.c_int), ("flags", C.c_int),
                ("pix_scale", C.c_float), ("device", C.c_int), ("arith", C.c_int),
                ("log_mult", C.c_float), ("power_gain", C.c_float),
                ("reserved", C.c_int * 5)]


class FseaSweepConfig(C.Structure):
    _fields_ = [("fft_size", C.c_int), ("rows", C.c_int), ("gate_rows", C.c_int), ("gate_threshold", C.c_double),
                ("pix_scale", C.c_float), ("in_format", C.c_int), ("device", C.c_int), ("capacity", C.c_int),
                ("own_raw", C.c_int), ("own_images", C.c_int), ("pixel_arith", C.c_int), ("reserved", C.c_int * 3)]


class FseaChainConfig(C.Structure):
    _fields_ = [("fft_size", C.c_int), ("block_samples", C.c_int), ("shifter", C.c_int), ("freq_offset", C.c_int),
                ("shifter_sample_rate", C.c_int), ("filter", C.c_int), ("taps", C.c_int), ("coefficients", C.POINTER(C.c_double)),
                ("device", C.c_int), ("reserved", C.c_int * 4)]


class FseaError(RuntimeError):
    def __init__(self, code: int, what: str, detail: str = ""):
        self.code = code
        super().__init__(f"{what}: fsea error {code}" + (f" ({detail})" if detail else ""))


def lib_path() -> str:
    return _LIB_PATH


def build(verbose: bool = False) -> str:
    """Compile libfsea.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j", str(os.cpu_count() or 4)]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0 or verbose:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building libfsea.so failed")
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load libfsea.so.  Fails loudly when the extension is missing: there is
    no Python / CPU implementation of the hot path behind this package."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise FileNotFoundError(
            f"{_LIB_PATH} not built -- run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the CUDA extension is required; there is no CPU fallback)")
    L = C.CDLL(_LIB_PATH)
    vp, i, l = C.c_void_p, C.c_int, C.c_long
    P = C.POINTER(FseaConfig)
    L.fsea_device_count.restype = i
    L.fsea_strerror.argtypes = [i]; L.fsea_strerror.restype = C.c_char_p
    L.fsea_last_cuda_error.restype = C.c_char_p
    L.fsea_config_default.argtypes = [P, i]; L.fsea_config_default.restype = None
    L.fsea_plan_create.argtypes = [P, C.POINTER(vp)]; L.fsea_plan_create.restype = i
    L.fsea_plan_destroy.argtypes = [vp]; L.fsea_plan_destroy.restype = i
    L.fsea_input_bytes.argtypes = [vp, l]; L.fsea_input_bytes.restype = C.c_size_t
    L.fsea_output_bytes.argtypes = [vp, l]; L.fsea_output_bytes.restype = C.c_size_t
    L.fsea_frames_in.argtypes = [vp, C.c_size_t]; L.fsea_frames_in.restype = l
    L.fsea_exec_device.argtypes = [vp, vp, vp, l, vp]; L.fsea_exec_device.restype = i
    L.fsea_exec_host.argtypes = [vp, vp, vp, l]; L.fsea_exec_host.restype = i
    L.fsea_exec_file.argtypes = [vp, C.c_char_p, l, l, vp]; L.fsea_exec_file.restype = i
    L.fsea_fir_iq_host.argtypes = [i, vp, i, vp, vp, l, vp]; L.fsea_fir_iq_host.restype = i
    L.fsea_plan_launch_count.argtypes = [vp]; L.fsea_plan_launch_count.restype = l
    L.fsea_plan_kernel_name.argtypes = [vp]; L.fsea_plan_kernel_name.restype = C.c_char_p
    L.fsea_gate_mean_host.argtypes = [vp, vp, l, C.POINTER(C.c_double)]; L.fsea_gate_mean_host.restype = i
    L.fsea_stitch_host.argtypes = [i, vp, i, i, i, i, vp]; L.fsea_stitch_host.restype = i
    L.fsea_stitch_device.argtypes = [vp, i, i, i, i, vp, vp]; L.fsea_stitch_device.restype = i
    L.fsea_stitch_sharded_device.argtypes = [vp, i, i, i, i, i, i, vp, vp]; L.fsea_stitch_sharded_device.restype = i
    L.fsea_exec_device_strided.argtypes = [vp, vp, vp, l, l, vp]; L.fsea_exec_device_strided.restype = i
    L.fsea_copy_only_host.argtypes = [vp, vp, vp, l]; L.fsea_copy_only_host.restype = i
    L.fsea_host_bind_near_device.argtypes = [i]; L.fsea_host_bind_near_device.restype = i
    L.fsea_chain_create.argtypes = [C.POINTER(FseaChainConfig), C.POINTER(vp)]; L.fsea_chain_create.restype = i
    L.fsea_chain_destroy.argtypes = [vp]; L.fsea_chain_destroy.restype = i
    L.fsea_chain_process_u8.argtypes = [vp, vp, vp]; L.fsea_chain_process_u8.restype = i
    L.fsea_chain_launch_count.argtypes = [vp]; L.fsea_chain_launch_count.restype = l
    L.fsea_history_put_device.argtypes = [vp, i, vp, i, vp]; L.fsea_history_put_device.restype = i
    L.fsea_history_shift_device.argtypes = [vp, vp, i, i, i, vp]; L.fsea_history_shift_device.restype = i
    L.fsea_ipc_alloc.argtypes = [i, C.c_size_t, C.POINTER(vp), C.c_char_p]; L.fsea_ipc_alloc.restype = i
    L.fsea_ipc_open.argtypes = [i, C.c_char_p, C.POINTER(vp)]; L.fsea_ipc_open.restype = i
    L.fsea_ipc_close.argtypes = [vp]; L.fsea_ipc_close.restype = i
    L.fsea_ipc_free.argtypes = [vp]; L.fsea_ipc_free.restype = i
    PS = C.POINTER(FseaSweepConfig)
    L.fsea_sweep_create.argtypes = [PS, C.POINTER(vp)]; L.fsea_sweep_create.restype = i
    L.fsea_sweep_destroy.argtypes = [vp]; L.fsea_sweep_destroy.restype = i
    L.fsea_sweep_submit_host.argtypes = [vp, i, vp]; L.fsea_sweep_submit_host.restype = i
    L.fsea_sweep_submit_device.argtypes = [vp, i, vp]; L.fsea_sweep_submit_device.restype = i
    L.fsea_sweep_submit_synth.argtypes = [vp, i, i, i]; L.fsea_sweep_submit_synth.restype = i
    L.fsea_sweep_gate.argtypes = [vp, i, vp, vp]; L.fsea_sweep_gate.restype = i
    L.fsea_sweep_gate_begin.argtypes = [vp, i]; L.fsea_sweep_gate_begin.restype = i
    L.fsea_sweep_gate_end.argtypes = [vp, i, vp, vp]; L.fsea_sweep_gate_end.restype = i
    L.fsea_sweep_render.argtypes = [vp, i, vp, l, l, l]; L.fsea_sweep_render.restype = i
    L.fsea_sweep_sync.argtypes = [vp]; L.fsea_sweep_sync.restype = i
    L.fsea_sweep_read_images.argtypes = [vp, i, i, vp]; L.fsea_sweep_read_images.restype = i
    L.fsea_sweep_read_raw.argtypes = [vp, i, vp]; L.fsea_sweep_read_raw.restype = i
    L.fsea_sweep_images.argtypes = [vp]; L.fsea_sweep_images.restype = vp
    L.fsea_sweep_raw.argtypes = [vp, i]; L.fsea_sweep_raw.restype = vp
    L.fsea_sweep_stream.argtypes = [vp]; L.fsea_sweep_stream.restype = vp
    L.fsea_sweep_launch_count.argtypes = [vp]; L.fsea_sweep_launch_count.restype = l
    _lib = L
    return L


def device_count() -> int:
    return int(lib().fsea_device_count())


def _check(rc: int, what: str) -> None:
    if rc != OK:
        L = lib()
        detail = L.fsea_strerror(rc).decode()
        if rc == ERR_CUDA:
            detail += ": " + L.fsea_last_cuda_error().decode()
        raise FseaError(rc, what, detail)


def stitch(images: np.ndarray, width_step: int, device: int = 0) -> np.ndarray:
    """Max-blend sweep images (n_images, rows, width) side by side (c/fft-stitch-broad.c)."""
    images = np.ascontiguousarray(images, dtype=np.uint8)
    n_images, rows, width = images.shape
    out = np.empty((rows, width + (n_images - 1) * width_step), dtype=np.uint8)
    _check(lib().fsea_stitch_host(device, images.ctypes.data_as(C.c_void_p), n_images, rows, width, width_step,
                                  out.ctypes.data_as(C.c_void_p)), "fsea_stitch_host")
    return out


_IN_NP = {IN_U8: np.uint8, IN_S8: np.uint8, IN_F32C: np.float32, IN_F64C: np.float64}


class Plan:
    """A batched IQ -> spectrum transform of fixed geometry (fsea_plan)."""

    def __init__(self, fft_size: int, hop: int = 0, in_format: int = IN_U8, window: int = WIN_NONE,
                 out_mode: int = OUT_MAG, flags: int = 0, pix_scale: float = 10.0, device: int = 0,
                 arith: int = ARITH_AUTO, log_mult: float = 10.0, power_gain: float = 1.0):
        L = lib()
        cfg = FseaConfig()
        L.fsea_config_default(C.byref(cfg), fft_size)
        cfg.hop = hop or fft_size
        cfg.in_format, cfg.window, cfg.out_mode, cfg.flags = in_format, window, out_mode, flags
        cfg.pix_scale, cfg.device, cfg.arith = pix_scale, device, arith
        cfg.log_mult, cfg.power_gain = log_mult, power_gain
        self.cfg = cfg
        self.n = fft_size
        self.hop = cfg.hop
        self.in_format, self.out_mode, self.arith = in_format, out_mode, arith
        self._h = C.c_void_p()
        _check(L.fsea_plan_create(C.byref(cfg), C.byref(self._h)), "fsea_plan_create")

    # -- geometry
    def input_bytes(self, n_frames: int) -> int:
        return int(lib().fsea_input_bytes(self._h, n_frames))

    def output_bytes(self, n_frames: int) -> int:
        return int(lib().fsea_output_bytes(self._h, n_frames))

    def frames_in(self, n_samples: int) -> int:
        return int(lib().fsea_frames_in(self._h, n_samples))

    @property
    def out_dtype(self):
        if self.out_mode == OUT_PIX:
            return np.uint8
        return np.float64 if (self.in_format == IN_F64C or self.arith == ARITH_F64) else np.float32

    @property
    def kernel_name(self) -> str:
        return lib().fsea_plan_kernel_name(self._h).decode()

    @property
    def launches(self) -> int:
        return int(lib().fsea_plan_launch_count(self._h))

    # -- execution
    def exec_device(self, d_in: int, d_out: int, n_frames: int, stream: int = 0) -> None:
        """Device pointers (ints); asynchronous on `stream` (a cudaStream_t handle)."""
        _check(lib().fsea_exec_device(self._h, C.c_void_p(d_in), C.c_void_p(d_out), n_frames,
                                      C.c_void_p(stream)), "fsea_exec_device")

    def exec_device_strided(self, d_in: int, d_out: int, out_row_stride_bytes: int, n_frames: int, stream: int = 0) -> None:
        """As exec_device with the output rows `out_row_stride_bytes` apart (rows inside a wider picture)."""
        _check(lib().fsea_exec_device_strided(self._h, C.c_void_p(d_in), C.c_void_p(d_out), out_row_stride_bytes, n_frames,
                                              C.c_void_p(stream)), "fsea_exec_device_strided")

    def copy_only_host_ptr(self, h_in: int, h_out: int, n_frames: int) -> None:
        """exec_host_ptr's copies without the kernel (the PCIe ceiling of the same job)."""
        _check(lib().fsea_copy_only_host(self._h, C.c_void_p(h_in), C.c_void_p(h_out), n_frames), "fsea_copy_only_host")

    def exec_host(self, data: np.ndarray, n_frames: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
        """Host buffers through the pinned H2D -> kernel -> D2H pipeline."""
        data = np.ascontiguousarray(data, dtype=_IN_NP[self.in_format]).ravel()
        per = 2
        if n_frames is None:
            n_frames = self.frames_in(data.size // per)
        need = self.input_bytes(n_frames)
        if data.nbytes < need:
            raise ValueError(f"input too short: {data.nbytes} < {need} bytes for {n_frames} frames")
        if out is None:
            out = np.empty((n_frames, self.n), dtype=self.out_dtype)
        assert out.flags.c_contiguous and out.nbytes >= self.output_bytes(n_frames)
        _check(lib().fsea_exec_host(self._h, data.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                    n_frames), "fsea_exec_host")
        return out

    def exec_file(self, path: str, n_frames: int, byte_offset: int = 0, out: np.ndarray | None = None) -> np.ndarray:
        """A capture file (raw 8-bit IQ as the SDR delivered it, c/rfcap.c / c/reader.c) through the same pipeline."""
        if out is None:
            out = np.empty((n_frames, self.n), dtype=self.out_dtype)
        assert out.flags.c_contiguous and out.nbytes >= self.output_bytes(n_frames)
        _check(lib().fsea_exec_file(self._h, os.fsencode(path), byte_offset, n_frames, out.ctypes.data_as(C.c_void_p)),
               "fsea_exec_file")
        return out

    def exec_host_ptr(self, h_in: int, h_out: int, n_frames: int) -> None:
        """Raw host pointers (e.g. pinned torch tensors: copied from / to directly)."""
        _check(lib().fsea_exec_host(self._h, C.c_void_p(h_in), C.c_void_p(h_out), n_frames), "fsea_exec_host")

    def gate_mean(self, data: np.ndarray, rows: int) -> float:
        data = np.ascontiguousarray(data, dtype=_IN_NP[self.in_format]).ravel()
        if data.nbytes < self.input_bytes(rows):
            raise ValueError(f"input too short: {data.nbytes} < {self.input_bytes(rows)} bytes for {rows} rows")
        m = C.c_double()
        _check(lib().fsea_gate_mean_host(self._h, data.ctypes.data_as(C.c_void_p), rows, C.byref(m)),
               "fsea_gate_mean_host")
        return float(m.value)

    def close(self) -> None:
        if self._h:
            lib().fsea_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def host_bind_near_device(device: int = 0) -> int:
    """Pin this thread (and the pinned buffers it allocates next) to the GPU's NUMA node; CPUs bound, 0 = unknown topology."""
    return int(lib().fsea_host_bind_near_device(device))


class Sweep:
    """The device-resident sweep engine of one GPU (fsea_sweep): raw IQ of `capacity` centre frequencies
    -> gate decisions -> u8 pixel rows, nothing crossing the host in between (c/fft-batch-broad.c:56-129)."""

    def __init__(self, fft_size: int, rows: int, capacity: int, gate_rows: int = 100, gate_threshold: float = 1.1,
                 pix_scale: float = 5.0, in_format: int = IN_S8, device: int = 0, own_raw: bool = True, own_images: bool = False,
                 pixel_arith: int = ARITH_AUTO):
        cfg = FseaSweepConfig(fft_size=fft_size, rows=rows, gate_rows=gate_rows, gate_threshold=gate_threshold,
                              pix_scale=pix_scale, in_format=in_format, device=device, capacity=capacity,
                              own_raw=int(own_raw), own_images=int(own_images), pixel_arith=pixel_arith)
        self.n, self.rows, self.capacity, self.device = fft_size, rows, capacity, device
        self._h = C.c_void_p()
        _check(lib().fsea_sweep_create(C.byref(cfg), C.byref(self._h)), "fsea_sweep_create")

    def submit_host(self, slot: int, raw: np.ndarray) -> None:
        raw = np.ascontiguousarray(raw, dtype=np.uint8).ravel()
        if raw.nbytes < 2 * self.n * self.rows:
            raise ValueError("raw IQ of a frequency must hold rows * N complex samples")
        _check(lib().fsea_sweep_submit_host(self._h, slot, raw.ctypes.data_as(C.c_void_p)), "fsea_sweep_submit_host")
        lib().fsea_sweep_sync(self._h)                        # (pageable source: the copy must be over before `raw` may go away)

    def submit_device(self, slot: int, d_raw: int) -> None:
        _check(lib().fsea_sweep_submit_device(self._h, slot, C.c_void_p(d_raw)), "fsea_sweep_submit_device")

    def submit_synth(self, slot: int, freq_idx: int, quiet_every: int = 10) -> None:
        _check(lib().fsea_sweep_submit_synth(self._h, slot, freq_idx, quiet_every), "fsea_sweep_submit_synth")

    def gate(self, n_slots: int):
        means = np.zeros(n_slots, dtype=np.float64)
        keep = np.zeros(n_slots, dtype=np.uint8)
        if n_slots > 0:
            _check(lib().fsea_sweep_gate(self._h, n_slots, means.ctypes.data_as(C.c_void_p), keep.ctypes.data_as(C.c_void_p)), "fsea_sweep_gate")
        return means, keep.astype(bool)

    def render(self, n_slots: int, d_canvas: int = 0, canvas_stride: int = 0, col0: int = 0, col_step: int = 0) -> None:
        if n_slots > 0:
            _check(lib().fsea_sweep_render(self._h, n_slots, C.c_void_p(d_canvas) if d_canvas else None, canvas_stride, col0, col_step),
                   "fsea_sweep_render")

    def sync(self) -> None:
        _check(lib().fsea_sweep_sync(self._h), "fsea_sweep_sync")

    def read_images(self, first: int, count: int) -> np.ndarray:
        out = np.empty((count, self.rows, self.n), dtype=np.uint8)
        _check(lib().fsea_sweep_read_images(self._h, first, count, out.ctypes.data_as(C.c_void_p)), "fsea_sweep_read_images")
        return out

    def read_raw(self, slot: int) -> np.ndarray:
        out = np.empty(2 * self.rows * self.n, dtype=np.uint8)
        _check(lib().fsea_sweep_read_raw(self._h, slot, out.ctypes.data_as(C.c_void_p)), "fsea_sweep_read_raw")
        return out

    def raw_ptr(self, slot: int) -> int:
        return int(lib().fsea_sweep_raw(self._h, slot) or 0)

    @property
    def stream(self) -> int:
        return int(lib().fsea_sweep_stream(self._h) or 0)

    @property
    def launches(self) -> int:
        return int(lib().fsea_sweep_launch_count(self._h))

    def close(self) -> None:
        if self._h:
            lib().fsea_sweep_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SharedDeviceBuffer:
    """Device memory another process of the box can store into (fsea_ipc_*): owner side `alloc`, peer side `open`."""

    def __init__(self, ptr: int, nbytes: int, owner: bool, handle: bytes = b""):
        self.ptr, self.nbytes, self.owner, self.handle = ptr, nbytes, owner, handle

    @classmethod
    def alloc(cls, device: int, nbytes: int) -> "SharedDeviceBuffer":
        p = C.c_void_p()
        h = C.create_string_buffer(64)
        _check(lib().fsea_ipc_alloc(device, nbytes, C.byref(p), h), "fsea_ipc_alloc")
        return cls(int(p.value), nbytes, True, h.raw)

    @classmethod
    def open(cls, device: int, handle: bytes, nbytes: int) -> "SharedDeviceBuffer":
        p = C.c_void_p()
        _check(lib().fsea_ipc_open(device, C.create_string_buffer(handle, 64), C.byref(p)), "fsea_ipc_open")
        return cls(int(p.value), nbytes, False)

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2, "strides": None}

    def close(self) -> None:
        if self.ptr:
            (lib().fsea_ipc_free if self.owner else lib().fsea_ipc_close)(C.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Chain:
    """Device block -> [frequency shifter] -> [IQ low-pass] -> nrf_fft row, fused on the device (fsea_chain_*)."""

    def __init__(self, fft_size: int, block_samples: int = 131072, shifter=None, fir_coefficients=None, device: int = 0):
        """shifter: (freq_offset, sample_rate) or None; fir_coefficients: the taps (nrf_fir_get_low_pass_coefficients) or None"""
        cfg = FseaChainConfig(fft_size=fft_size, block_samples=block_samples, device=device)
        self._coef = None
        if shifter is not None:
            cfg.shifter, cfg.freq_offset, cfg.shifter_sample_rate = 1, int(shifter[0]), int(shifter[1])
        if fir_coefficients is not None:
            self._coef = np.ascontiguousarray(fir_coefficients, dtype=np.float64)
            cfg.filter, cfg.taps = 1, int(self._coef.size)
            cfg.coefficients = self._coef.ctypes.data_as(C.POINTER(C.c_double))
        self.n, self.block_samples = fft_size, block_samples
        self._h = C.c_void_p()
        _check(lib().fsea_chain_create(C.byref(cfg), C.byref(self._h)), "fsea_chain_create")

    def process(self, block_u8: np.ndarray) -> np.ndarray:
        block_u8 = np.ascontiguousarray(block_u8, dtype=np.uint8).ravel()
        if block_u8.size < 2 * self.block_samples:
            raise ValueError("a device block holds 2 * block_samples bytes")
        row = np.empty(self.n, dtype=np.float64)
        _check(lib().fsea_chain_process_u8(self._h, block_u8.ctypes.data_as(C.c_void_p), row.ctypes.data_as(C.c_void_p)), "fsea_chain_process_u8")
        return row

    @property
    def launches(self) -> int:
        return int(lib().fsea_chain_launch_count(self._h))

    def close(self) -> None:
        if self._h:
            lib().fsea_chain_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
