"""Whole-codec (GPU transform + host brotli) throughput vs batch / threads / frame count."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from fusion_power_video_b200 import host, synth
W, H, bits, shift = 1280, 800, 12, 4
P = W * H
base = synth.plasma_frames(64, W, H, bits=bits, seed=5).reshape(64, -1)
ncpu = os.cpu_count()
host.time_encode(base[:2], W, H, shift, threads=4, batch=2)
for n, batch, threads in [(64, 32, ncpu), (256, 32, ncpu), (256, 8, ncpu), (256, 4, ncpu), (512, 8, ncpu), (1024, 8, ncpu), (1024, 16, ncpu), (1024, 8, ncpu - 1), (1024, 8, ncpu // 2)]:
    fr = np.ascontiguousarray(np.tile(base, (n // 64, 1)))
    t, size = host.time_encode(fr, W, H, shift, threads=threads, batch=batch)
    print(f"n={n} batch={batch} threads={threads}: {t*1e3:.0f} ms  {n*P/t/1e6:.0f} MP/s  {n/t:.0f} fps", flush=True)
try:
    from oracle_binding import Ref
    ref = Ref()
    for n in (64, 256):
        fr = np.ascontiguousarray(np.tile(base, (n // 64, 1)))
        t, size = ref.time_encode(fr, W, H, shift, 0, fr[0], ncpu)
        print(f"reference n={n} threads={ncpu}: {t*1e3:.0f} ms {n*P/t/1e6:.0f} MP/s")
except Exception as e:
    print("no ref", e)
