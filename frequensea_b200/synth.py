"""SURVEY.md 8(d) synthetic IQ sources, drawn on a torch device (measurement tooling for bench.py).

The reference ships captures, not generators; BASELINE's synthetic configurations are defined in
SURVEY.md 8(d).  tests/conftest.py holds the numpy statements of the same formulas (`synth_u8`,
`synth_s8`); these torch versions exist because 2^28 samples per step cannot come from a host
generator in bench time.  Same signal model, torch's generator instead of numpy's: the bytes differ,
the statistics do not, and every parity gate runs the oracle on the bytes that were actually timed.
"""
from __future__ import annotations

import math


def synth_u8_cfg3(n_samples: int, device, seed: int = 20260925, chunk: int = 1 << 24):
    """cfg3: offset-binary uint8 IQ (RTL-SDR convention): three tones at f/fs = 0.0371, -0.2113, 0.3307 with
    amplitudes 1, 0.5, 0.1, scaled 127.5 * 0.25 around 127.5, plus N(0, 20^2) noise; clip, round."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty(2 * n_samples, dtype=torch.uint8, device=device)
    for s0 in range(0, n_samples, chunk):
        m = min(chunk, n_samples - s0)
        t = torch.arange(s0, s0 + m, dtype=torch.float64, device=device)
        re = torch.zeros(m, dtype=torch.float64, device=device)
        im = torch.zeros(m, dtype=torch.float64, device=device)
        for f, a in ((0.0371, 1.0), (-0.2113, 0.5), (0.3307, 0.1)):
            ph = 2.0 * math.pi * torch.remainder(f * t, 1.0)
            re += a * torch.cos(ph)
            im += a * torch.sin(ph)
        noise = torch.randn((2, m), dtype=torch.float32, device=device, generator=g) * 20.0
        i = 127.5 + 127.5 * 0.25 * re.float() + noise[0]
        q = 127.5 + 127.5 * 0.25 * im.float() + noise[1]
        v = out[2 * s0: 2 * (s0 + m)].view(m, 2)
        v[:, 0] = torch.clamp(torch.round(i), 0, 255).to(torch.uint8)
        v[:, 1] = torch.clamp(torch.round(q), 0, 255).to(torch.uint8)
    return out


def synth_s8_cfg5(n_samples: int, device, seed: int = 7, chunk: int = 1 << 24):
    """cfg5: two's-complement int8 IQ (HackRF convention, returned as uint8 bytes): sigma-24 noise + a chirp
    of amplitude 40 sweeping 0.05 .. 0.25 fs over the stream."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty(2 * n_samples, dtype=torch.uint8, device=device)
    for s0 in range(0, n_samples, chunk):
        m = min(chunk, n_samples - s0)
        t = torch.arange(s0, s0 + m, dtype=torch.float64, device=device)
        cyc = 0.05 * t + 0.2 * t * t / (2.0 * max(n_samples, 1))
        ph = 2.0 * math.pi * torch.remainder(cyc, 1.0)
        noise = torch.randn((2, m), dtype=torch.float32, device=device, generator=g) * 24.0
        i = 40.0 * torch.cos(ph).float() + noise[0]
        q = 40.0 * torch.sin(ph).float() + noise[1]
        v = out[2 * s0: 2 * (s0 + m)].view(m, 2)
        v[:, 0] = torch.clamp(torch.round(i), -128, 127).to(torch.int8).view(torch.uint8)
        v[:, 1] = torch.clamp(torch.round(q), -128, 127).to(torch.int8).view(torch.uint8)
    return out
