#!/usr/bin/env python
"""Per-role cycle accounting of k_decode_pair (profiling build: make LIBDIR=../lib_prof EXTRA=-DFPV_PAIR_PROF).

    FPV_B200_LIB=$PWD/fusion_power_video_b200/lib_prof/libfpv_b200.so python scripts/gpu_pair_prof.py [c2 c1 c3]

Prints, per workload, the average cycles per row a chain / IO / helper warp spends in each phase and how
many repair rounds a row takes.  The counters are clock() deltas summed over all warps (g_pair_prof)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import fusion_power_video_b200 as fpv  # noqa: E402
from fusion_power_video_b200 import synth  # noqa: E402

WL = {"c2": (1280, 800, 12, 4), "c1": (1024, 1024, 16, 0), "c3": (2048, 2048, 16, 0)}


def main():
    names = sys.argv[1:] or ["c2", "c1", "c3"]
    L = fpv.lib()
    if not hasattr(L, "fpv_debug_pair_prof"):
        raise SystemExit("not a -DFPV_PAIR_PROF build: set FPV_B200_LIB")
    L.fpv_debug_pair_prof.argtypes = [C.c_void_p]
    out = {}
    for name in names:
        W, H, bits, shift = WL[name]
        P = W * H
        F = 1184 if name != "c3" else 592
        dev = torch.device("cuda", 0)
        frames = synth.plasma_frames_torch(F, W, H, bits=bits, seed=1, device=dev).reshape(F, P)
        hi = torch.empty((F, P), dtype=torch.uint8, device=dev)
        lo = torch.empty((F, P), dtype=torch.uint8, device=dev)
        pv = torch.empty((F, P // 16), dtype=torch.uint8, device=dev)
        fl = torch.empty(F, dtype=torch.uint8, device=dev)
        o = torch.empty((F, P), dtype=torch.int16, device=dev)
        ctx = fpv.Context(W, H, shift, False, max_batch=F)
        ctx.set_delta_raw_device(frames[0].data_ptr())
        ctx.encode_device(frames.data_ptr(), F, fl.data_ptr(), hi.data_ptr(), lo.data_ptr(), pv.data_ptr())
        for _ in range(2):
            ctx.decode_device(hi.data_ptr(), lo.data_ptr(), fl.data_ptr(), F, o.data_ptr(), options=fpv.DEC_UNEXTRACT)
        torch.cuda.synchronize()
        buf = (C.c_ulonglong * 16)()
        L.fpv_debug_pair_prof(buf)   # clear
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.decode_device(hi.data_ptr(), lo.data_ptr(), fl.data_ptr(), F, o.data_ptr(), options=fpv.DEC_UNEXTRACT)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        L.fpv_debug_pair_prof(buf)
        v = [int(x) for x in buf]
        rows = max(v[15], 1)          # chain-warp rows
        io_rows = rows                # one IO and one helper warp per chain warp
        res = {
            "ms": ms, "frames": F, "chain_rows": rows,
            "chain_cycles_per_row": {k: v[i] / rows for i, k in enumerate(["load_c", "pass0", "pass1", "repair", "store", "barrier_wait"])},
            "repair_rounds_per_row": v[6] / rows,
            "io_cycles_per_row": {k: v[8 + i] / io_rows for i, k in enumerate(["post_row", "tma_store_read_wait", "barrier_wait"])},
            "helper_cycles_per_row": {k: v[12 + i] / io_rows for i, k in enumerate(["issue", "pre_row", "barrier_wait"])},
        }
        res["row_cycles"] = sum(res["chain_cycles_per_row"].values())
        out[name] = res
        print(name, json.dumps(res))
        del frames, hi, lo, pv, fl, o, ctx
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "pair_prof.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
