#!/usr/bin/env python
"""PCIe ceilings of the box the bench runs on: pinned host <-> device copies, one direction at a time and both at once
(the e2e leg moves 2 B/px in and 2.06 B/px out concurrently, so the bidirectional figure is its bound)."""
import json
import sys
import time

import torch


def rate(fn, nbytes, reps=8):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


def main():
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d()
        d2h()

    res = {"chunk_mb": mb, "h2d_gbs": rate(h2d, n), "d2h_gbs": rate(d2h, n)}
    b = rate(both, n)          # bytes per direction per second
    res["bidir_each_gbs"] = b
    print(json.dumps(res))


if __name__ == "__main__":
    main()
