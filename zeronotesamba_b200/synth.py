"""Deterministic synthetic 16 kHz drum / other stem pairs (no Spleeter, no files, no network).

Stand-ins for the stems the reference cuts out of FMA with Spleeter
(/root/reference/zeroNoteSamba/pretext.py:30-86, fma_loader.py:91-150): a *drums* stem (click
train of decaying noise bursts plus a low decaying sine) and an *other* stem (harmonic notes
that change every beat, slow tremolo), both over a -60 dBFS white-noise floor so that every VQT
bin stays above the float32 precision floor.  Same ``(seed, clip_idx)`` -> same samples on every
machine (numpy Generator PCG64 only).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

SAMPLE_RATE = 16000


def stem_pair(clip_idx: int, seconds: float = 10.0, seed: int = 1234, sr: int = SAMPLE_RATE,
              n_samples: int | None = None) -> Tuple[np.ndarray, np.ndarray]:
    """Return ``(drums, other)`` float32 arrays of ``int(seconds * sr)`` samples, peak <= 0.9."""
    rng = np.random.default_rng(seed + clip_idx)
    n = int(round(seconds * sr)) if n_samples is None else int(n_samples)
    t = np.arange(n, dtype=np.float64) / sr
    bpm = rng.uniform(60.0, 180.0)
    beat = 60.0 / bpm
    phase0 = rng.uniform(0.0, beat)
    hits = np.arange(phase0, n / sr, beat)

    drums = np.zeros(n, dtype=np.float64)
    for h in hits:
        tau = rng.uniform(0.010, 0.080)
        f_lo = rng.uniform(50.0, 120.0)
        i0 = int(h * sr)
        ln = min(n - i0, int(6 * tau * sr))
        if ln <= 0:
            continue
        tt = np.arange(ln) / sr
        env = np.exp(-tt / tau)
        drums[i0:i0 + ln] += env * (0.6 * rng.standard_normal(ln) + 0.8 * np.sin(2 * np.pi * f_lo * tt))

    other = np.zeros(n, dtype=np.float64)
    edges = np.concatenate([[0.0], hits, [n / sr]])
    for a, b in zip(edges[:-1], edges[1:]):
        i0, i1 = int(a * sr), int(b * sr)
        if i1 <= i0:
            continue
        tt = t[i0:i1] - a
        seg = np.zeros(i1 - i0)
        for _ in range(int(rng.integers(3, 7))):
            f0 = 55.0 * 2.0 ** rng.uniform(0.0, 4.0)  # 55..880 Hz
            ph = rng.uniform(0, 2 * np.pi)
            for hnum in range(1, 9):
                if f0 * hnum < 0.45 * sr:
                    seg += np.sin(2 * np.pi * f0 * hnum * tt + ph * hnum) / hnum
        fade = np.minimum(1.0, np.minimum(tt, (b - a) - tt) / 0.005)
        other[i0:i1] += seg * fade
    other *= 1.0 + 0.3 * np.sin(2 * np.pi * rng.uniform(0.2, 2.0) * t)

    noise_rms = 10.0 ** (-60.0 / 20.0)
    out = []
    for sig in (drums, other):
        peak = np.max(np.abs(sig))
        if peak > 0:
            sig = sig * (0.9 / peak) * 0.98
        sig = sig + noise_rms * rng.standard_normal(n)
        sig = np.clip(sig, -0.9, 0.9)
        out.append(sig.astype(np.float32))
    return out[0], out[1]


def stem_batch(first_clip: int, count: int, seconds: float = 10.0, seed: int = 1234) -> Tuple[np.ndarray, np.ndarray]:
    """``(drums[count, N], other[count, N])`` for clips ``first_clip .. first_clip + count - 1``."""
    pairs = [stem_pair(first_clip + i, seconds, seed) for i in range(count)]
    return np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])


def cfg2_batch(device, n_clips: int = 256, seconds: float = 30.0, n_base: int = 8):
    """BASELINE.json configs[1] input: ``n_clips`` distinct ``seconds``-long 16 kHz clips as a CUDA fp32 tensor.

    ``n_base`` synthetic stems (drums / other alternating) are tiled and made distinct by seeded -80 dBFS device noise
    (torch.Generator(device).manual_seed(1234)), so bench.py and the parity test see the same batch."""
    import torch
    base = np.stack([stem_pair(i, seconds)[i % 2] for i in range(n_base)])
    y = torch.from_numpy(base).to(device).repeat((n_clips + n_base - 1) // n_base, 1)[:n_clips].contiguous()
    gen = torch.Generator(device=device)
    gen.manual_seed(1234)
    y += 1e-4 * torch.randn(y.shape, device=device, generator=gen)
    return y
