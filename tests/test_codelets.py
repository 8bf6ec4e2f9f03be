"""CPU suite: the generated register codelets (fsea_codelets.cuh) compiled for
the host in float and double against numpy's DFT."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("codelet") / "codelet_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "codelet_host.cpp")], check=True)
    L = C.CDLL(so)
    L.codelet_run.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.codelet_run_raw.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return L


@pytest.mark.parametrize("R", [2, 4, 8, 16, 32])
def test_codelet(harness, R):
    rng = np.random.default_rng(R)
    for is_double, tol in ((1, 1e-14), (0, 5e-7)):
        x = rng.standard_normal(R) + 1j * rng.standard_normal(R)
        re, im = x.real.copy(), x.imag.copy()
        harness.codelet_run(R, is_double, re.ctypes.data, im.ctypes.data)
        ref = np.fft.fft(x)
        assert np.abs(re + 1j * im - ref).max() <= tol * np.abs(ref).max() * R


@pytest.mark.parametrize("R", [8, 16, 32])
def test_raw_codelet(harness, R):
    """First stage consumes 32768 + byte directly; exact for every byte pattern."""
    rng = np.random.default_rng(R)
    for is_double, tol in ((1, 1e-13), (0, 5e-7)):
        b = rng.integers(0, 256, (2, R)).astype(np.float64)
        re, im = 32768.0 + b[0], 32768.0 + b[1]
        harness.codelet_run_raw(R, is_double, re.ctypes.data, im.ctypes.data)
        ref = np.fft.fft((b[0] - 128) + 1j * (b[1] - 128))
        assert np.abs(re + 1j * im - ref).max() <= tol * max(1.0, np.abs(ref).max()) * R


def test_generated_header_is_current():
    out = subprocess.run(["python", os.path.join(ROOT, "tools", "gen_codelets.py")], check=True, stdout=subprocess.PIPE, text=True).stdout
    assert out == open(os.path.join(ROOT, "frequensea_b200", "csrc", "fsea_codelets.cuh")).read()
