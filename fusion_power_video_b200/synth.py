"""Seeded synthetic "noisy plasma" frames.

Smooth bright blobs (2-D Gaussians drifting slowly with the frame index) on a
static background of about 2 % full scale, shot-like noise (sigma ~ sqrt of the
signal) plus a small read noise, clipped to [0, 2^bits - 1] and stored as
native little-endian uint16 -- NOT left-aligned, so 12-bit data is encoded with
shift 4.  The numpy generator is used by the tests and the CPU baseline; the
torch generator builds bench-sized batches directly in HBM.
"""
from __future__ import annotations

import numpy as np


def _blob_params(seed):
    rng = np.random.default_rng(seed)
    k = 3
    return dict(
        cx=rng.uniform(0.25, 0.75, k), cy=rng.uniform(0.25, 0.75, k),
        sx=rng.uniform(0.08, 0.25, k), sy=rng.uniform(0.08, 0.25, k),
        amp=rng.uniform(0.15, 0.6, k), vx=rng.uniform(-2e-3, 2e-3, k), vy=rng.uniform(-2e-3, 2e-3, k),
    )


def plasma_frames(n, xsize, ysize, bits=16, seed=0, first=0):
    """uint16 [n, ysize, xsize]; frame t = first + index."""
    full = float((1 << bits) - 1)
    bp = _blob_params(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, ysize, dtype=np.float32), np.linspace(0, 1, xsize, dtype=np.float32), indexing="ij")
    out = np.empty((n, ysize, xsize), np.uint16)
    for i in range(n):
        t = first + i
        sig = np.full((ysize, xsize), 0.02, np.float32)
        for k in range(len(bp["cx"])):
            cx = bp["cx"][k] + bp["vx"][k] * t
            cy = bp["cy"][k] + bp["vy"][k] * t
            sig += bp["amp"][k] * np.exp(-(((xx - cx) / bp["sx"][k]) ** 2 + ((yy - cy) / bp["sy"][k]) ** 2) * 0.5).astype(np.float32)
        sig *= full
        rng = np.random.default_rng([seed, t, 7919])
        noise = rng.standard_normal((ysize, xsize), dtype=np.float32)
        read = rng.standard_normal((ysize, xsize), dtype=np.float32)
        # shot noise in "photo-electrons": gain so that sigma stays a few hundred counts at 16 bit
        gain = full / 4096.0
        val = sig + noise * np.sqrt(np.maximum(sig, 0) * gain) + read * (full * 2e-4 + 1.0)
        out[i] = np.clip(np.rint(val), 0, full).astype(np.uint16)
    return out


def plasma_frames_torch(n, xsize, ysize, bits=16, seed=0, first=0, device="cuda"):
    """Same model generated on the device (not bit-identical to the numpy one)."""
    import torch

    full = float((1 << bits) - 1)
    bp = _blob_params(seed)
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1000003 + first)
    yy, xx = torch.meshgrid(
        torch.linspace(0, 1, ysize, device=device), torch.linspace(0, 1, xsize, device=device), indexing="ij"
    )
    out = torch.empty((n, ysize, xsize), dtype=torch.uint16, device=device)
    gain = full / 4096.0
    for i in range(n):
        t = first + i
        sig = torch.full((ysize, xsize), 0.02, device=device)
        for k in range(len(bp["cx"])):
            cx = float(bp["cx"][k] + bp["vx"][k] * t)
            cy = float(bp["cy"][k] + bp["vy"][k] * t)
            sig = sig + float(bp["amp"][k]) * torch.exp(
                -(((xx - cx) / float(bp["sx"][k])) ** 2 + ((yy - cy) / float(bp["sy"][k])) ** 2) * 0.5
            )
        sig = sig * full
        noise = torch.randn((ysize, xsize), device=device, generator=g)
        read = torch.randn((ysize, xsize), device=device, generator=g)
        val = sig + noise * torch.sqrt(torch.clamp(sig, min=0) * gain) + read * (full * 2e-4 + 1.0)
        out[i] = torch.clamp(torch.round(val), 0, full).to(torch.int32).to(torch.uint16)
    return out
