"""The kernel SOURCES on the CPU fiber machine (tests/emu) against the oracle -- no GPU needed.

tests/emu compiles frequensea_b200/csrc/*.cu with g++ (-DFSEA_EMU) against a fake
cuda_runtime.h: every CUDA thread is a fiber, barriers / shuffles / cluster barriers / cp.async
have their CUDA meaning, packed f32x2 arithmetic is two scalar ops with the device's
contraction rules.  It checks indexing, synchronisation and arithmetic of the kernels a
reviewer reads; it says nothing about speed.  The emulated library is TEST INFRASTRUCTURE: it
is never loaded by the product (frequensea_b200 resolves its own libfsea.so; FSEA_TEST_LIB is read by
tests/conftest.py and tests/emu/check.py, which bind the package to the emulated build explicitly).  The -m gpu suite repeats every one of
these comparisons on the device through the real library.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu_lib():
    r = subprocess.run(["make", "-C", EMU, "-j8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return os.path.join(EMU, "libfsea_emu.so")


def run_check(emu_lib, sizes, **extra_env):
    env = dict(os.environ, FSEA_TEST_LIB=emu_lib, **extra_env)
    r = subprocess.run([sys.executable, os.path.join(EMU, "check.py")] + [str(n) for n in sizes],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def verify(res):
    for r in res:
        tag = f"N={r['n']} {r['kernel']}"
        assert r["mag_patch"] < 1e-5, tag            # the north star's 1e-5 relative bound (rms-normalised)
        assert r["patch_dup"], tag                   # out[N/2] == out[N/2-1], src/nrf.c:621-622
        assert r["logmag_db"] < 2e-3, tag            # dB, bins >= 1e-3 rms
        assert r["dc_bin_db"] < 1e-3, tag            # the analytic N/2 (1+i) DC term of uncentred u8/256
        assert r["scope_err"] < 1.4e-3, tag          # 7 log10(): the 2e-3 dB bound scaled by 7/10
        assert r["pix_max"] <= 1 and r["pix_frac"] < 2e-4, tag   # trunc() of a float: <= 1 LSB on < 2e-4 of the pixels (measured <= 3e-5)


def test_emulated_single_launch_sizes(emu_lib):
    """N = 16 .. 8192: fsea_fft_kernel in every default arrangement (pairs, warp tickets, frame per unit)."""
    verify(run_check(emu_lib, [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192]))


def test_emulated_alternative_arrangements(emu_lib):
    """The arrangements FSEA_ARITH_AUTO does not pick: frame pairs (2), frame per unit (3) at 2048 ... 8192,
    scalar fp32 (1)."""
    res = run_check(emu_lib, ["2048:2", "4096:2", "4096:3", "8192:2", "1024:1", "4096:1"])
    assert "64x64" not in res[2]["kernel"] and "frame/unit" in res[2]["kernel"]
    verify(res)


def test_emulated_four_step_sizes(emu_lib):
    """N = 16384 .. 65536: the defaults -- the warp-specialised block-per-frame kernel (16384), its cluster form with
    the intermediate in distributed shared memory and TMA-fed tiles (32768), and the L2 cluster kernel (65536:
    sub-spectra, cluster barrier, comb) incl. slot reuse."""
    res = run_check(emu_lib, [16384, 32768, 65536])
    assert "ws4<" in res[0]["kernel"] and "wsc<" in res[1]["kernel"] and "fft4<" in res[2]["kernel"]
    verify(res)


def test_emulated_warp_specialised_and_cluster_alternatives(emu_lib):
    """fsea_ws4_kernel at every length it exists for (FSEA_WS=1: 4, 2 and 1 frames per block iteration; named
    barriers between transform and comb warps, register gathers, frames past the end of the batch), and
    the cluster kernel at 16384 (FSEA_FOUR_ALT=9), which the default no longer picks."""
    res = run_check(emu_lib, [4096, 8192, 16384, 32768, 65536], FSEA_WS="1")
    assert all("ws4<" in r["kernel"] for r in res[:3]) and all("wsc<" in r["kernel"] for r in res[3:])     # mbarriers, TMA boxes, stores into peer blocks
    verify(res)
    res = run_check(emu_lib, [16384, 32768], FSEA_FOUR_ALT="9")
    assert all("fft4<" in r["kernel"] for r in res)
    verify(res)


def test_emulated_iq_filter_bit_exact(emu_lib):
    """nrf_iq_filter_* (SURVEY 8(f) rank 3): the device FIR kernel, on the fiber machine, equals the
    compiled reference bit for bit -- odd, even and short tap counts, state carried across blocks."""
    from oracle import orc
    if not orc.have_ref():
        pytest.skip("oracle/_ref/libnrf_ref.so not built")
    env = dict(os.environ, FSEA_TEST_LIB=emu_lib)
    r = subprocess.run([sys.executable, os.path.join(EMU, "check.py"), "iq_filter"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert json.loads(r.stdout.strip().splitlines()[-1])[0]["max_abs_diff"] == 0.0


def test_emulated_capture_file_feed(emu_lib):
    """fsea_exec_file: pread-fed pipeline == host-buffer pipeline, byte offsets,
    overlapping frames, several slots in flight; short / missing / misaligned files are refused."""
    env = dict(os.environ, FSEA_TEST_LIB=emu_lib)
    r = subprocess.run([sys.executable, os.path.join(EMU, "check.py"), "file"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert json.loads(r.stdout.strip().splitlines()[-1])[0]["identical"] is True


def test_gpu_suite_on_the_fiber_machine(emu_lib, tmp_path):
    """The -m gpu parity tests themselves (C ABI, operator mirror, sweep job, the reference's Lua
    scripts on tools/_ref/fsea-lua) with the emulated library in libfsea.so's place: the test logic
    and every kernel source are exercised before a GPU minute is spent.  Left out: the two tests that
    allocate device memory through torch.  On the B200 box the same tests run against the real library."""
    os.symlink(emu_lib, tmp_path / "libfsea.so")                     # for the Lua host's DT_NEEDED libfsea.so
    env = dict(os.environ, FSEA_TEST_LIB=emu_lib, LD_LIBRARY_PATH=str(tmp_path) + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"),
                        os.path.join(ROOT, "tests", "test_gpu_lua.py"), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider",
                        "-k", "not many_frames and not large_batch and not sweep_job and not mgpu"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
