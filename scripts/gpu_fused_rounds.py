#!/usr/bin/env python
"""Histogram of repair rounds per row of k_decode_fused (profiling build:
make LIBDIR=../lib_prof EXTRA=-DFPV_FUSED_PROF ../lib_prof/libfpv_b200.so;
FPV_B200_LIB=$PWD/fusion_power_video_b200/lib_prof/libfpv_b200.so python scripts/gpu_fused_rounds.py)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fusion_power_video_b200 as fpv  # noqa: E402
from fusion_power_video_b200 import synth  # noqa: E402

WL = {"c2": (1280, 800, 12, 4, 592), "c1": (1024, 1024, 16, 0, 592), "c3": (2048, 2048, 16, 0, 296)}


def main():
    L = fpv.lib()
    if not hasattr(L, "fpv_debug_fused_rounds"):
        raise SystemExit("not a -DFPV_FUSED_PROF build: set FPV_B200_LIB")
    out = {}
    for name, (W, H, bits, shift, F) in WL.items():
        P = W * H
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
        torch.cuda.synchronize()
        buf = (C.c_ulonglong * 8)()
        L.fpv_debug_fused_rounds(buf)   # clear
        ctx.decode_device(hi.data_ptr(), lo.data_ptr(), fl.data_ptr(), F, o.data_ptr(), options=fpv.DEC_UNEXTRACT)
        torch.cuda.synchronize()
        L.fpv_debug_fused_rounds(buf)
        v = [int(x) for x in buf]
        rows = sum(v)
        out[name] = {"geometry": f"{W}x{H} {bits}-bit", "frames": F, "rows_of_a_pair_or_split_frame": rows,
                     "share_of_rows_by_repair_rounds_0_to_7plus": [round(x / max(rows, 1), 5) for x in v],
                     "mean_rounds_per_row": round(sum(i * x for i, x in enumerate(v)) / max(rows, 1), 4),
                     "exact": bool(torch.equal(o.view(torch.uint16), frames))}
        print(name, json.dumps(out[name]))
        del frames, hi, lo, pv, fl, o, ctx
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fused_rounds.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
