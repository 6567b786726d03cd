#!/usr/bin/env python
"""GPU entropy decoder on a batch of full-size frames (for ncu captures and a wall-clock figure):
frames -> fpv_encode_stream (GPU coder with chunk directories) -> fpv_decode_coded -> raw frames."""
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
import huffcoder_ref as href


def main():
    W, H, bits, shift, n = 1280, 800, 12, 4, int(sys.argv[1]) if len(sys.argv) > 1 else 64
    P = W * H
    frames = synth.plasma_frames(n, W, H, bits=bits, seed=1).reshape(n, -1)
    with fpv.Context(W, H, shift, False, max_batch=n) as ctx:
        ctx.set_delta_raw(frames[0])
        flags, chunks = ctx.encode_stream(frames)
        blob, table = bytearray(), []
        for f, ch in enumerate(chunks):
            bp1 = struct.unpack_from("<I", ch, 5)[0]
            core = ch[10 + bp1 - 1:]
            pos = 1
            for plane in ([1, 0] if not core[0] & 4 else [0]):
                offs, length = href.scan_plane(core[pos:], P)
                table += [(len(blob) + pos + o, f, plane, k) for k, o in enumerate(offs)]
                pos += length
            blob += core
            blob += bytes(-len(blob) % 16)
        blob = bytes(blob)
        for _ in range(2):
            raw = ctx.decode_coded(blob, table, flags, fpv.DEC_UNEXTRACT)
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            raw = ctx.decode_coded(blob, table, flags, fpv.DEC_UNEXTRACT)
        dt = (time.perf_counter() - t0) / reps
        print(f"fpv_decode_coded: {n} frames, {len(blob) / 1e6:.1f} MB coded, {len(table)} chunks, {dt * 1e3:.2f} ms per call "
              f"(incl. the Python-side table build? no: call only; H2D + entropy decode + inverse + D2H), "
              f"{n * P * 2 / dt / 1e9:.1f} GB/s raw, exact {bool(np.array_equal(raw, frames))}")


if __name__ == "__main__":
    main()
