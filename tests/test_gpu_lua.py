"""GPU suite: the reference's own lua/fft*.lua scripts, UNCHANGED (compiled to Lua
bytecode by tools/Makefile), driven by the headless host
tools/_ref/fsea-lua against libfsea.so; the texture payload each script hands to
ngl_texture_update is compared with the oracle on the same capture blocks."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import orc
from conftest import BLOCK

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "tools", "_ref", "fsea-lua")
TREE = os.path.join(ROOT, "tools", "_ref", "tree")

needs_host = pytest.mark.skipif(not os.path.exists(HOST) or not os.path.exists(os.path.join(TREE, "lua", "fft.lua")),
                                reason="tools/_ref (script host + compiled scripts) not built: needs /root/reference at build time")


def run_script(name, frames, tmp_path, extra=()):
    dump = str(tmp_path / f"{os.path.basename(name)}.dump")
    path = name if os.path.isabs(name) else f"../lua/{name}.lua"
    cmd = [HOST, path, "--frames", str(frames), "--step", "--dump", dump, *extra]
    res = subprocess.run(cmd, cwd=os.path.join(TREE, "run"), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    out = []
    with open(dump, "rb") as f:
        while True:
            hdr = f.read(12)
            if len(hdr) < 12:
                break
            frame, w, h = struct.unpack("<iii", hdr)
            out.append((frame, np.frombuffer(f.read(4 * w * h), dtype=np.float32).reshape(h, w)))
    return out, res.stdout


def rel(got, ref):
    return float(np.max(np.abs(got.astype(np.float64) - ref)) / np.sqrt(np.mean(ref ** 2)))


@needs_host
@pytest.mark.parametrize("script,n,hist", [("fft", 1024, 1024), ("fft-sea", 128, 512), ("fft-sea-auto", 128, 512), ("fft-sea-sick", 128, 128)])
def test_fft_scripts_run_unchanged(script, n, hist, iq, tmp_path):
    frames = 4
    updates, stdout = run_script(script, frames, tmp_path)
    assert len(updates) == frames, stdout
    blocks = [orc.device_unpack(iq["rf200big4"][b * BLOCK:(b + 1) * BLOCK]) for b in range(4)]
    rows = [orc.nrf_fft_frame_u8(b, n) for b in blocks]
    for f, (frame, tex) in enumerate(updates):
        assert frame == f and tex.shape == (hist, n)
        for r in range(f + 1):                       # row 0 newest (src/nrf.c:617-631)
            assert rel(tex[r], rows[f - r]) < 1e-5, (script, f, r)
        assert not tex[f + 1:].any()


@needs_host
def test_fft_shifted_script(iq, tmp_path):
    """lua/fft-shifted.lua: device -> nrf_freq_shifter (shift 0) -> nrf_fft on a float64 buffer."""
    updates, _ = run_script("fft-shifted", 2, tmp_path)
    for f, (_, tex) in enumerate(updates):
        u8 = orc.device_unpack(iq["rf200big4"][f * BLOCK:(f + 1) * BLOCK])
        x = u8[:2048].astype(np.float64) / 256.0 + 0.5          # shifter output at zero offset (src/nrf.c:856-857)
        assert rel(tex[0], orc.nrf_fft_frame_f64(x, 1024)) < 1e-6


@needs_host
def test_retune_key_shifts_history(iq, tmp_path):
    """on_key -> keys_frequency_handler -> nrf_fft_shift (lua/_keys.lua:146-169): +0.1 MHz at 5 MS/s
    scrolls every history row by round(1024 / 50) = 20 bins."""
    updates, _ = run_script("fft", 3, tmp_path, extra=("--key", "2:KEY_RIGHT:0"))
    before, after = updates[1][1], updates[2][1]
    s = int(round(1024 / ((5e6 / 1e6) / 0.1)))
    assert s == 20
    # rows 1.. of the last frame are the previous rows 0.., shifted left by s and zero-filled
    assert np.array_equal(after[1:3, : 1024 - s], before[0:2, s:])
    assert not after[1:3, 1024 - s:].any()


@needs_host
def test_filter_chain_script(iq, tmp_path):
    """tests/lua/filtered-fft.lua (ours): device -> nrf_iq_filter -> nrf_fft, the operator in front of the
    path of lua/iq-tex-filtered.lua:43-46 -- against the compiled reference's own chain on the same blocks."""
    if not orc.have_ref():
        pytest.skip("oracle/_ref/libnrf_ref.so not built")
    import ctypes as C
    frames = 3
    updates, stdout = run_script(os.path.join(ROOT, "tests", "lua", "filtered-fft.lua"), frames, tmp_path)
    assert len(updates) == frames, stdout
    blocks = [orc.device_unpack(iq["rf200big4"][b * BLOCK:(b + 1) * BLOCK]) for b in range(frames)]
    filtered = orc.ref_iq_filter_blocks(blocks, 5000000, 200000, 51)
    R = orc.ref()
    rfft = R.nrf_fft_new(1024, 8)
    for f, (frame, tex) in enumerate(updates):
        nb = R.nut_buffer_new_f64(filtered[f].size // 2, 2, filtered[f].ctypes.data_as(C.c_void_p))
        R.nrf_fft_process(rfft, nb)
        R.nut_buffer_free(nb)
        o = R.nrf_fft_get_buffer(rfft)
        ref = orc.nut_to_numpy(o).reshape(8, 1024)
        R.nut_buffer_free(o)
        assert tex.shape == (8, 1024)
        for r in range(f + 1):
            assert rel(tex[r], ref[r]) < 1e-6, (f, r)          # float64 transform, float32 texture
    R.nrf_fft_free(rfft)
