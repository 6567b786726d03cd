#!/usr/bin/env python
"""Randomised parity sweep (not part of the pytest suite): random geometries, split modes and contents through
encode / decode / the GPU entropy coder, every result compared with the oracle (tests/oracle_binding.py) or the
CPU restatement of the bitstream.  Usage: python scripts/gpu_fuzz.py [seconds] [seed]"""
import os
import struct
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import fusion_power_video_b200 as fpv
from fusion_power_video_b200 import synth
from oracle_binding import Oracle
import huffcoder_ref as href


def content(rng, kind, n, W, H, bits):
    P = W * H
    if kind == 0:
        return synth.plasma_frames(n, W, H, bits=bits, seed=int(rng.integers(1 << 30))).reshape(n, -1)
    if kind == 1:
        return rng.integers(0, 1 << bits, (n, P)).astype(np.uint16)
    if kind == 2:
        base = rng.integers(0, 1 << bits, P).astype(np.uint16)
        return np.stack([(base + rng.integers(0, 3, P)).astype(np.uint16) & ((1 << bits) - 1) for _ in range(n)])
    if kind == 3:
        v = rng.integers(0, 1 << bits)
        f = np.full((n, P), v, np.uint16)
        f[:, :: max(1, P // 7)] ^= 1
        return f
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    return np.stack([(((xx * (3 + k) + yy * (5 + 2 * k)) & ((1 << bits) - 1))).astype(np.uint16).reshape(-1) for k in range(n)])


def entropy_case(rng):
    """Random symbol distributions straight into the entropy coder (fpv_entropy_device), vs the CPU restatement."""
    import torch

    W = 16 * int(rng.integers(4, 120))
    H = 4 * int(rng.integers(1, 40))
    n = int(rng.integers(1, 4))
    P, PP = W * H, (W // 4) * (H // 4)

    def plane(size):
        k = int(rng.integers(6))
        if k == 0:
            return np.full(size, rng.integers(256), np.uint8)
        if k == 1:
            return rng.choice(rng.integers(0, 256, int(rng.integers(2, 6))), size).astype(np.uint8)
        if k == 2:
            return np.minimum(rng.geometric(float(rng.uniform(0.01, 0.9)), size) - 1, 255).astype(np.uint8)
        if k == 3:
            return rng.integers(0, 256, size).astype(np.uint8)
        if k == 4:   # exponentially spaced counts: deep trees
            w = 1.6 ** -np.arange(int(rng.integers(8, 40)))
            return rng.choice(len(w), size, p=w / w.sum()).astype(np.uint8)
        return (np.cumsum(rng.integers(-2, 3, size)) & 255).astype(np.uint8)

    high = np.stack([plane(P) for _ in range(n)])
    low = np.stack([plane(P) for _ in range(n)])
    prev = np.stack([plane(PP) for _ in range(n)])
    flags = rng.integers(0, 8, n).astype(np.uint8)
    dev = torch.device("cuda", 0)
    th, tl, tp, tf = (torch.from_numpy(a).to(dev) for a in (high, low, prev, flags))
    with fpv.Context(W, H, 0, False, max_batch=n) as ctx:
        cap = ctx.stream_bound(n)
        out = torch.zeros(cap, dtype=torch.uint8, device=dev)
        off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        ctx.entropy_device(tf.data_ptr(), th.data_ptr(), tl.data_ptr(), tp.data_ptr(), n, out.data_ptr(), cap, off.data_ptr())
        torch.cuda.synchronize()
    off, out = off.cpu().numpy(), out.cpu().numpy()
    for i in range(n):
        fl = int(flags[i])
        bp = href.encode_plane(prev[i])
        core = bytes([fl]) + (b"" if fl & 4 else href.encode_plane(low[i])) + href.encode_plane(high[i])
        exp = struct.pack("<IBIB", 10 + len(bp) + len(core), 0, len(bp) + 1, (fl & 2) | 4) + bp + core
        assert out[off[i]:off[i + 1]].tobytes() == exp, f"entropy_device W={W} H={H} frame {i} flags {fl}"
        assert href.brotli_decode(href.encode_plane(high[i]), P) == high[i].tobytes()


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    oracle = Oracle()
    t0 = time.time()
    cases = fails = 0
    while time.time() - t0 < budget:
        W = int(rng.choice([4 * int(rng.integers(1, 80)), 8 * int(rng.integers(8, 330)), 32 * int(rng.integers(41, 81)),
                            256 * int(rng.integers(1, 9)), 16 * int(rng.integers(4, 81))]))
        H = 4 * int(rng.integers(1, 24)) if rng.integers(4) else 4 * int(rng.integers(24, 90))   # tall: several bands
        very_tall = rng.integers(10) == 0           # H >= 2048: every band count and ring wrap of the full-size frames
        if very_tall:
            H = 4 * int(rng.integers(512, 640))
        shift, be = [(0, 0), (4, 0), (8, 0), (0, 1), (3, 1), (8, 1), (6, 0), (12, 0)][int(rng.integers(8))]
        bits = 16 - shift if shift <= 8 else 4
        n = int(rng.integers(1, 7)) if not very_tall else int(rng.integers(1, 4))
        kind = int(rng.integers(5))
        frames = content(rng, kind, n, W, H, bits)
        delta = frames[0] if rng.integers(2) else content(rng, int(rng.integers(5)), 1, W, H, bits)[0]
        if be:
            frames, delta = frames.byteswap(), delta.byteswap()
        P, PP = W * H, (W // 4) * (H // 4)
        tag = f"W={W} H={H} shift={shift} be={be} n={n} kind={kind}"
        try:
            if cases % 7 == 3:
                entropy_case(rng)
            with fpv.Context(W, H, shift, bool(be), max_batch=4) as ctx:
                ctx.set_delta_raw(delta)
                flags, high, low, prev = ctx.encode(frames)
                out = ctx.decode(high, low, flags)
                raw = ctx.decode(high, low, flags, fpv.DEC_UNEXTRACT)
                sflags, chunks = ctx.encode_stream(frames[:4]) if not very_tall else (None, None)
            dimg = oracle.delta_image(delta, shift, be)
            for i in range(n):
                fl, h, l, p = oracle.predict(frames[i], W, H, shift, be, delta)
                assert fl == int(flags[i]), f"flags {fl} vs {int(flags[i])}"
                assert np.array_equal(h, high[i]) and np.array_equal(p, prev[i]), "high / preview plane"
                if low is not None and not (fl & 4):
                    assert np.array_equal(l, low[i]), "low plane"
                exp = oracle.inverse(high[i], None if (fl & 4) or low is None else low[i], dimg, W, H, fl)
                assert np.array_equal(out[i], exp), "decoded image"
                assert np.array_equal(raw[i].view(np.uint8), oracle.unextract(exp, shift, be)), "unextract"
            if W % 4 == 0 and cases % 5 == 0 and not very_tall:
                # host layer: Encoder (both entropy stages) -> StreamingDecoder must reproduce the raw file
                from fusion_power_video_b200 import host
                for ge in (False, True):
                    st = host.encode_stream(frames, W, H, shift, bool(be), threads=3, batch=int(rng.integers(1, 5)), delta=delta,
                                            gpu_entropy=ge)
                    back = host.decode_stream(st, n + 1, W, H, block=int(rng.choice([0, 4096, 1 << 20])), batch=3,
                                              raw_shift=shift, big_endian=bool(be))
                    expect = frames if shift == 0 else None
                    dec_img = host.decode_stream(st, n + 1, W, H, batch=2)
                    assert dec_img.shape[0] == n and np.array_equal(dec_img, out), f"host decode (gpu_entropy={ge})"
                    assert back.shape[0] == n and np.array_equal(back, raw), f"host raw decode (gpu_entropy={ge})"
            for i in range(min(n, 4) if not very_tall else 0):   # (the Python restatement of the coder is slow on 5 MP planes)
                fl = int(flags[i])
                bp = href.encode_plane(prev[i])
                core = bytes([fl]) + (b"" if (fl & 4) or low is None else href.encode_plane(low[i])) + href.encode_plane(high[i])
                exp_chunk = struct.pack("<IBIB", 10 + len(bp) + len(core), 0, len(bp) + 1, (fl & 2) | 4) + bp + core
                assert chunks[i] == exp_chunk, "entropy coder chunk"
        except Exception as e:  # noqa: BLE001
            fails += 1
            print("FAIL", tag, "->", repr(e)[:200], flush=True)
        cases += 1
    print(f"fuzz: {cases} cases, {fails} failures, seed {seed}, {time.time() - t0:.0f} s")
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
