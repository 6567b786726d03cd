"""Seeded input cases shared by the golden-vector generator and the tests.

Each case is (name, W, H, shift, big_endian, frames uint16 [n, H, W],
delta_raw uint16 [H, W] or None).  Inputs are regenerated from the seed; the
golden files additionally store them so that a numpy RNG change cannot move
the pin silently.
"""
from __future__ import annotations

import numpy as np

from fusion_power_video_b200 import synth


def _byteswap(a):
    return a.byteswap()


def make_case(name):
    rng = np.random.default_rng(abs(hash(name)) % (2**31) if False else sum(ord(c) * (i + 1) for i, c in enumerate(name)))
    if name == "plasma16_le0":
        f = synth.plasma_frames(4, 64, 48, bits=16, seed=11)
        return dict(W=64, H=48, shift=0, be=0, frames=f, delta=f[0])
    if name == "plasma12_le4":
        f = synth.plasma_frames(4, 80, 40, bits=12, seed=12)
        return dict(W=80, H=40, shift=4, be=0, frames=f, delta=f[0])
    if name == "plasma8_le8":
        f = synth.plasma_frames(3, 64, 32, bits=8, seed=13)
        return dict(W=64, H=32, shift=8, be=0, frames=f, delta=f[0])
    if name == "plasma16_be0":
        f = _byteswap(synth.plasma_frames(3, 64, 48, bits=16, seed=14))
        return dict(W=64, H=48, shift=0, be=1, frames=f, delta=f[0])
    if name == "plasma12_be4":
        f = _byteswap(synth.plasma_frames(3, 72, 36, bits=12, seed=15))
        return dict(W=72, H=36, shift=4, be=1, frames=f, delta=f[0])
    if name == "plasma8_be8":
        f = _byteswap(synth.plasma_frames(3, 64, 32, bits=8, seed=16))
        return dict(W=64, H=32, shift=8, be=1, frames=f, delta=f[0])
    if name == "plasma4_le12":
        f = synth.plasma_frames(3, 64, 32, bits=4, seed=17)
        return dict(W=64, H=32, shift=12, be=0, frames=f, delta=f[0])
    if name == "noise16_le0":  # CG should lose on white noise; delta frame is noise too
        f = rng.integers(0, 65536, (3, 40, 56), dtype=np.uint16)
        d = rng.integers(0, 65536, (40, 56), dtype=np.uint16)
        return dict(W=56, H=40, shift=0, be=0, frames=f, delta=d)
    if name == "be3_fullrange":  # exercises the swapped-endian generic branch with non-zero top bits
        f = rng.integers(0, 65536, (3, 32, 48), dtype=np.uint16)
        return dict(W=48, H=32, shift=3, be=1, frames=f, delta=f[1])
    if name == "le5_fullrange":
        f = rng.integers(0, 65536, (3, 32, 48), dtype=np.uint16)
        return dict(W=48, H=32, shift=5, be=0, frames=f, delta=f[1])
    if name == "constant_high":  # single-bin histogram: delta NOT chosen (EstimateEntropy == 0)
        f = (np.full((3, 24, 32), 0x1200, np.uint16) + rng.integers(0, 256, (3, 24, 32), dtype=np.uint16)).astype(np.uint16)
        return dict(W=32, H=24, shift=0, be=0, frames=f, delta=f[0])
    if name == "all_zero":  # NO_LOW_BYTES without delta
        f = np.zeros((2, 16, 16), np.uint16)
        return dict(W=16, H=16, shift=0, be=0, frames=f, delta=None)
    if name == "zero_low_bytes":  # NO_LOW_BYTES, delta frame shares the property (round-trippable)
        f = (synth.plasma_frames(3, 32, 24, bits=8, seed=18).astype(np.uint16) << 8).astype(np.uint16)
        return dict(W=32, H=24, shift=0, be=0, frames=f, delta=f[0])
    if name == "no_delta_frame":  # Predict(EMPTY)
        f = synth.plasma_frames(3, 64, 32, bits=16, seed=19)
        return dict(W=64, H=32, shift=0, be=0, frames=f, delta=None)
    if name == "frame_equals_delta":
        f = synth.plasma_frames(2, 48, 32, bits=16, seed=20)
        f[1] = f[0]
        return dict(W=48, H=32, shift=0, be=0, frames=f, delta=f[0])
    if name == "ramp":  # smooth gradient: CG wins clearly, many wrap-arounds in the high byte
        yy, xx = np.meshgrid(np.arange(40), np.arange(64), indexing="ij")
        f = np.stack([((xx * 997 + yy * 1361 + t * 4099) & 0xFFFF).astype(np.uint16) for t in range(3)])
        return dict(W=64, H=40, shift=0, be=0, frames=f, delta=f[0])
    if name == "narrow_w4":  # minimum legal width: every pixel sits next to a row wrap
        f = rng.integers(0, 65536, (3, 64, 4), dtype=np.uint16)
        return dict(W=4, H=64, shift=0, be=0, frames=f, delta=f[2])
    if name == "w12_h8":  # W % 8 != 0: generic (non-TMA) path
        f = synth.plasma_frames(3, 12, 8, bits=16, seed=21)
        return dict(W=12, H=8, shift=0, be=0, frames=f, delta=f[0])
    if name == "wide_w520":  # more than two 256-column strips, ragged last strip
        f = synth.plasma_frames(2, 520, 16, bits=16, seed=22)
        return dict(W=520, H=16, shift=0, be=0, frames=f, delta=f[0])
    raise KeyError(name)


CASE_NAMES = [
    "plasma16_le0", "plasma12_le4", "plasma8_le8", "plasma16_be0", "plasma12_be4", "plasma8_be8",
    "plasma4_le12", "noise16_le0", "be3_fullrange", "le5_fullrange", "constant_high", "all_zero",
    "zero_low_bytes", "no_delta_frame", "frame_equals_delta", "ramp", "narrow_w4", "w12_h8", "wide_w520",
]
