import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
BLOCK = 262144


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu() -> bool:
    try:
        import frequensea_b200 as fb
        return fb.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a device must fail loudly, not skip: the product has no CPU path.
    pass


@pytest.fixture(scope="session")
def iq():
    z = np.load(os.path.join(GOLDEN, "iq_fixtures.npz"))
    return {"rf433": z["rf433"], "rf200big4": z["rf200big4"]}


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(GOLDEN, "golden_ref.npz")))


@pytest.fixture(scope="session")
def fb():
    import frequensea_b200
    if not os.path.exists(frequensea_b200.lib_path()):
        frequensea_b200.build()
    return frequensea_b200


def synth_u8(n_samples: int, seed: int = 20260925) -> np.ndarray:
    """SURVEY.md 8(d) cfg3 generator: three tones + noise, offset-binary uint8 IQ."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64)
    sig = np.zeros(n_samples, dtype=np.complex128)
    for f, a in ((0.0371, 1.0), (-0.2113, 0.5), (0.3307, 0.1)):
        sig += a * np.exp(2j * np.pi * f * t)
    i = 127.5 + 127.5 * 0.25 * sig.real + rng.normal(0, 20, n_samples)
    q = 127.5 + 127.5 * 0.25 * sig.imag + rng.normal(0, 20, n_samples)
    out = np.empty(2 * n_samples, dtype=np.uint8)
    out[0::2] = np.clip(np.round(i), 0, 255)
    out[1::2] = np.clip(np.round(q), 0, 255)
    return out


def synth_s8(n_samples: int, seed: int = 7) -> np.ndarray:
    """SURVEY.md 8(d) cfg5 generator: sigma-24 noise + chirp, two's-complement int8 IQ (as uint8 bytes)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64)
    ph = 2 * np.pi * (0.05 * t + 0.2 * t * t / (2.0 * max(n_samples, 1)))
    i = 40 * np.cos(ph) + rng.normal(0, 24, n_samples)
    q = 40 * np.sin(ph) + rng.normal(0, 24, n_samples)
    out = np.empty(2 * n_samples, dtype=np.int8)
    out[0::2] = np.clip(np.round(i), -128, 127)
    out[1::2] = np.clip(np.round(q), -128, 127)
    return out.view(np.uint8)
