#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference
(oracle/_ref/libfpv_ref.so, built from /root/reference by oracle/Makefile).

Run from the repo root where /root/reference exists:
    make -C oracle all && python tests/golden/make_golden.py

The reference ships no golden vectors of its own; these
files are the pin for the oracle and, through it, for the CUDA path.  Each
case file stores the inputs too, so a change in numpy's RNG cannot move the
pin silently.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ctypes as C  # noqa: E402

from cases import CASE_NAMES, make_case  # noqa: E402
from oracle_binding import Ref, _p  # noqa: E402


def main():
    ref = Ref()
    for name in CASE_NAMES:
        c = make_case(name)
        W, H, shift, be = c["W"], c["H"], c["shift"], c["be"]
        frames = np.ascontiguousarray(c["frames"], dtype=np.uint16).reshape(-1, W * H)
        delta = None if c["delta"] is None else np.ascontiguousarray(c["delta"], dtype=np.uint16).reshape(-1)
        n = frames.shape[0]
        flags = np.zeros(n, np.uint8)
        high = np.zeros((n, W * H), np.uint8)
        low = np.zeros((n, W * H), np.uint8)
        preview = np.zeros((n, (W // 4) * (H // 4)), np.uint8)
        has_low = shift != 8
        for i in range(n):
            fl, h, l, p = ref.predict(frames[i], W, H, shift, be, delta)
            flags[i], high[i], preview[i] = fl, h, p
            assert (l is not None) == has_low
            if l is not None:
                low[i] = l
        out = dict(W=W, H=H, shift=shift, be=be, frames=frames, flags=flags, high=high, preview=preview,
                   has_delta=np.uint8(delta is not None))
        if has_low:
            out["low"] = low
        # split-only view (Frame ctor), used for the delta planes
        if delta is not None:
            out["delta"] = delta
            # whole stream through the reference encoder and both reference decoders
            stream = ref.encode_stream(frames, W, H, shift, be, delta, threads=2)
            stream1 = ref.encode_stream(frames, W, H, shift, be, delta, threads=0)
            assert np.array_equal(stream, stream1), "reference output depends on thread count?"
            nd, dec, wo, ho = ref.decode_stream(stream, n, W, H, block=4096)
            assert nd == n and wo == W and ho == H
            out["decoded"] = dec
            out["stream_sha256"] = np.frombuffer(hashlib.sha256(stream.tobytes()).digest(), np.uint8)
            out["stream_size"] = np.uint64(stream.size)
            # the delta frame as the decoders see it
            dsize = int(np.frombuffer(stream[8:12].tobytes(), "<u4")[0])
            ok, dimg = ref.decompress_image(None, stream[13:8 + dsize], W, H)
            assert ok
            out["delta_image"] = dimg
            raw = np.zeros((n, W * H * 2), np.uint8)
            for i in range(n):
                raw[i] = ref.unextract(dec[i], W, H, shift, be)
            out["unextracted"] = raw
            # random-access decoder agrees with the streaming one
            for i in range(n):
                ok, fr, pv, nf = ref.random_access_decode(stream, i, W, H)
                assert ok and nf == n and np.array_equal(fr, dec[i])
            ok, fr, pv, nf = ref.random_access_decode(stream, n - 1, W, H)
            out["last_preview_decoded"] = pv
        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **out)
        print(f"{name}: flags={flags.tolist()} W={W} H={H} shift={shift} be={be}")

    # EstimateEntropy known answers: assorted histogram shapes
    rng = np.random.default_rng(20201)
    hists = []
    for k in range(400):
        kind = k % 8
        h = np.zeros(256, np.uint64)
        if kind == 0:
            h[:] = rng.integers(0, 5000, 256)
        elif kind == 1:
            h[rng.integers(0, 256)] = rng.integers(1, 10**7)
        elif kind == 2:
            idx = rng.integers(0, 256, 3)
            h[idx] = rng.integers(1, 50, 3)
            h[rng.integers(0, 256)] += 69906
        elif kind == 3:
            h[:] = (rng.exponential(200, 256)).astype(np.uint64)
        elif kind == 4:
            h[rng.integers(0, 256, 20)] = rng.integers(1, 4, 20)
        elif kind == 5:
            pass  # all zero
        elif kind == 6:
            h[:] = rng.integers(0, 2, 256) * rng.integers(1, 300000, 256)
        else:
            h[:8] = rng.integers(1, 2**22, 8)
        hists.append(h)
    hists = np.stack(hists)
    ent = np.array([ref.estimate_entropy(h) for h in hists], np.uint64)
    # ClampedGradient, exhaustive
    table = np.zeros(1 << 24, np.uint8)
    ref.L.ref_cg_table.argtypes = [C.c_void_p]
    ref.L.ref_cg_table(_p(table))
    np.savez_compressed(
        os.path.join(HERE, "scalars.npz"), hists=hists, entropy=ent,
        cg_table_sha256=np.frombuffer(hashlib.sha256(table.tobytes()).digest(), np.uint8),
        cg_table_sample=table[:: 4099].copy(),
    )
    print("scalars: entropy range", ent.min(), ent.max())


if __name__ == "__main__":
    main()
