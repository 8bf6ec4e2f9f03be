"""frequensea_b200 -- B200-native IQ -> spectrum engine behind frequensea's
nrf_* / nut_buffer_* operator API.

The product is the C-ABI shared library ``libfsea.so`` (CUDA kernels for
sm_100a + the C/C++ host mirror of the reference's operator interface, see
``include/fsea.h``, ``include/nrf.h``, ``include/nut.h``).  This package is the
thin ctypes binding used by the tests and by ``bench.py``; PyTorch is only used
by callers for device memory, streams and ``torch.distributed`` plumbing.

There is no CPU fallback: if ``libfsea.so`` is missing this module raises at
import of :func:`lib`, and every compute call fails with
``FSEA_ERR_NO_DEVICE`` when no CUDA device is visible.
"""
from .binding import (  # noqa: F401
    FseaError, Plan, build, device_count, lib, lib_path, stitch,
    IN_U8, IN_S8, IN_F32C, IN_F64C, WIN_NONE, WIN_HANN,
    OUT_MAG, OUT_LOGMAG, OUT_PIX, FLAG_DC_PATCH,
    ARITH_AUTO, ARITH_F32, ARITH_F32X2, ARITH_F32X2_FRAME, ARITH_F32X2_RADIX64, ARITH_F64,
)
from . import nrf  # noqa: F401
from . import shard, sweep  # noqa: F401

__all__ = ["FseaError", "Plan", "build", "device_count", "lib", "lib_path", "nrf"]
