#!/usr/bin/env python
"""StreamingDecoder throughput on a GPU-coded and a brotli stream for several GpuOptions::batch values."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from fusion_power_video_b200 import host, synth

W, H, bits, shift, n = 1280, 800, 12, 4, 1024
frames = synth.plasma_frames(256, W, H, bits=bits, seed=1).reshape(256, -1)
frames = np.ascontiguousarray(np.tile(frames, (4, 1)))
for ge in (True, False):
    st = host.encode_stream(frames, W, H, shift, False, threads=16, batch=32, gpu_entropy=ge)
    for batch in (32, 64, 128):
        best = None
        for _ in range(3):
            cnt, sec, first = host.decode_stream(st, n, W, H, block=0, batch=batch, raw_shift=shift, return_time="both", keep=False)
            if best is None or sec < best:
                best, steady = sec, (n - batch) / (sec - first)
        print(f"gpu_entropy={ge} batch={batch}: {n * W * H * 2 / best / 1e9:.2f} GB/s raw, {n / best:.0f} fps whole call; "
              f"steady {steady:.0f} fps = {steady * W * H * 2 / 1e9:.1f} GB/s", flush=True)
