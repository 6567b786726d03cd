#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 pre-entropy transform.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N = 1 (BASELINE.json configs[1]): 12-bit 1280x800 high-speed-camera frames
stored as uint16 (shift 4), static delta frame = frame 0, one GPU holding a
contiguous range of the sequence resident in HBM.  A *step* is one pass of the
encode transform (Frame ctor + Frame::Predict of the reference) over that
range.  The line also carries `configs` (the same measurement on configs[0] and
configs[2] geometry) and `ingest` (configs[4]: paced real-time ingest).

N > 1 (BASELINE.json configs[2]): ONE 16-bit 2048x2048 sequence of 10 000 frames
cut into contiguous frame ranges, one rank / GPU each (sharding.frame_range), no
data-path collective.  The only thing that crosses GPUs is the delta frame: rank
0 splits it and every other rank copies the planes device to device through a
CUDA IPC handle (fpv_delta_ipc_export / _import).  A step is one pass over the
whole sequence (every rank over its shard, resident in HBM); the total work is
fixed, so "scaling" is "strong".  The run ends with a merged-stream check: a
subsample of the sequence is encoded shard by shard with fpvc::Encoder, the
shard streams are merged on rank 0 (sharding.merge_shards) and compared with
the stream ONE GPU writes for the same frames.

Prints ONE JSON line (rank 0).  `value` is device-resident raw-pixel GB/s
(2 bytes x pixels / s) over all GPUs; `e2e` is the same metric through the
host-buffer C-ABI calls (pinned H2D + kernels + D2H inside the timed region);
`roofline` is the fused encode kernel against the measured HBM peak;
`cpu_baseline` is the reference's own CPU code (oracle/_ref) on this box's
host cores.  `--impl reference` times only that CPU code.

Further entries of the line: `e2e.pcie_ceiling` (the box's own host <-> device
copy ceiling with every rank copying at once, measured in the same run) and
`e2e.frac_of_pcie_ceiling`; `decode` (inverse transform, one full wave of the
decode kernel, bit-exact round trip); `decode_e2e`, `e2e_stream`; `entropy`
(GPU entropy coder on resident planes); `stream` (whole codec: host brotli,
GPU entropy stage from pageable and from page-locked frames, the decoders on
both kinds of stream incl. the GPU entropy decoder, next to the reference's
Encoder / StreamingDecoder); `configs` (configs[0] / configs[2] geometries);
`ingest` (configs[4]: paced real-time ingest); `multi_gpu` (N > 1: delta halo
by CUDA IPC, merged-stream check).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (xsize, ysize, bits, shift, description)
    "c2": (1280, 800, 12, 4, "12-bit 1280x800 camera stream, static delta frame, device-resident (BASELINE configs[1])"),
    "c1": (1024, 1024, 16, 0, "16-bit 1024x1024 (BASELINE configs[0] geometry)"),
    "c3": (2048, 2048, 16, 0, "16-bit 2048x2048 x 10k frames, encode sharded by frame range across the GPUs with a one-frame (delta frame) halo (BASELINE configs[2])"),
    # experiments: is a lower roofline fraction a matter of geometry (power-of-two strides) or of content (16-bit noise)?
    "x1": (1024, 1000, 16, 0, "experiment: 16-bit 1024x1000 (frame size not a power of two)"),
    "x2": (1280, 800, 16, 0, "experiment: 16-bit content at the C2 geometry"),
    "x3": (1024, 1024, 12, 4, "experiment: 12-bit content at the C1 geometry"),
}
ENC_BYTES_PER_PX = 2 + 1 + 1 + 1.0 / 16  # raw in, high out, low out, preview out
DEC_BYTES_PER_PX = 1 + 1 + 2             # planes in, uint16 out


def ncu_traffic(which, algorithmic_bytes):
    """DRAM bytes per launch from the committed ncu --set full capture (profiles/traffic.json), scaled to this
    launch's size by the measured DRAM-bytes-per-algorithmic-byte ratio; None if there is no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return float(json.load(f)[which]["dram_bytes_per_algorithmic_byte"]) * algorithmic_bytes
    except Exception:
        return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed regions run."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, windows):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if not any(a <= t <= b + 0.06 for a, b in windows):
                continue
            p = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(p[0])); mx.append(float(p[1])); power.append(float(p[2]))
            except Exception:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}


def pin_rank_cpus(local, rank, world):
    """Gives every rank of a multi-GPU run its own slice of the host cores -- inside the GPU's NUMA node where the
    platform reports one (sysfs), else of all cores -- BEFORE any pinned buffer is allocated, so that first-touch puts
    the staging memory next to the threads that fill it and the ranks' brotli / copy threads do not migrate onto each
    other.  Returns the cpu list (also used as the rank's thread budget)."""
    cpus = sorted(os.sched_getaffinity(0))
    if world <= 1:
        return cpus
    node_cpus = None
    try:
        bus = subprocess.check_output(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                                      text=True).strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node >= 0:
            node_cpus = []
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                a, _, b = part.partition("-")
                node_cpus.extend(range(int(a), int(b or a) + 1))
            node_cpus = sorted(set(node_cpus) & set(cpus))
    except Exception:
        node_cpus = None
    per = max(1, len(cpus) // world)
    mine = cpus[(rank * per) % len(cpus):(rank * per) % len(cpus) + per]
    if node_cpus:
        # ranks that share a node split that node's cores among themselves
        mine = [c for c in mine if c in node_cpus] or node_cpus[:per]
    try:
        os.sched_setaffinity(0, mine)
    except Exception:
        return cpus
    return mine


def cpu_reference_rate(frames_np, W, H, shift, delta_np, budget_s, threads=None):
    """raw-pixel GB/s of the reference's CPU transform (Frame ctor + Predict) on `threads` host threads."""
    from oracle_binding import Oracle, Ref, ref_available

    threads = threads or os.cpu_count() or 1
    n = frames_np.shape[0]
    if ref_available():
        ref = Ref()
        best, t_used, reps = None, 0.0, 0
        while t_used < budget_s or reps < 1:
            t = ref.time_transform(frames_np, W, H, shift, 0, delta_np, threads)
            best = t if best is None else min(best, t)
            t_used += t
            reps += 1
            if reps >= 50:
                break
        return n * W * H * 2 / best / 1e9, "reference", threads, reps
    oracle = Oracle()
    t0 = time.perf_counter()
    for i in range(n):
        oracle.predict(frames_np[i], W, H, shift, 0, delta_np)
    t = time.perf_counter() - t0
    return n * W * H * 2 / t / 1e9, "port", 1, 1


def run_reference_arm(args, W, H, bits, shift, desc):
    """--impl reference: the reference's own CPU implementation of the path, all host threads."""
    from fusion_power_video_b200 import synth
    from oracle_binding import Ref, ref_available

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = max(threads, min(8 * threads, 64))
    frames = synth.plasma_frames(per_step, W, H, bits=bits, seed=1).reshape(per_step, -1)
    delta = frames[0].copy()
    if not ref_available():
        rate, kind, cores, _ = cpu_reference_rate(frames[:4], W, H, shift, delta, 1.0)
        ms = 4 * W * H * 2 / rate / 1e6
    else:
        ref = Ref()
        # keep (steps + warmup) bounded to a few minutes: shrink the per-step sample if needed
        t_probe = ref.time_transform(frames, W, H, shift, 0, delta, threads)
        budget = 150.0
        while per_step > threads and t_probe * (args.steps + args.warmup) > budget:
            per_step = max(threads, per_step // 2)
            frames = frames[:per_step]
            t_probe = ref.time_transform(frames, W, H, shift, 0, delta, threads)
        for _ in range(args.warmup):
            ref.time_transform(frames, W, H, shift, 0, delta, threads)
        total = 0.0
        for _ in range(args.steps):
            total += ref.time_transform(frames, W, H, shift, 0, delta, threads)
        ms = total / args.steps * 1e3
        rate = per_step * W * H * 2 / (ms * 1e-3) / 1e9
        kind, cores = "reference", threads
    sample = f"{per_step} frames {W}x{H} per step, Frame ctor + Frame::Predict, {cores} host threads over frames"
    line = {
        "impl": "reference", "metric": "encode_transform_raw_pixel_throughput", "value": rate, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "frames_per_s": per_step / (ms * 1e-3),
        "config": {"workload": desc, "xsize": W, "ysize": H, "bits": bits, "shift": shift, "frames_per_step": per_step},
        "cpu_baseline": {"value": rate, "unit": "GB/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def pcie_probe(torch, dist, world, dev, seconds=0.5, mb=64):
    """Host <-> device copy ceilings of THIS box with all `world` ranks copying at once (pinned memory, 64 MiB pieces):
    H2D alone, D2H alone, both directions at the same time.  The host-buffer legs (`e2e`, `decode_e2e`) move about
    2 B/px each way concurrently, so their ceiling in raw-pixel GB/s is the `bidir_each` figure."""
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d()
        d2h()

    out = {"piece_mb": mb, "seconds_per_mode": seconds}
    for name, fn in (("h2d", h2d), ("d2h", d2h), ("bidir_each", both)):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < seconds:
            for _ in range(4):
                fn()
            torch.cuda.synchronize()
            reps += 4
        rate = n * reps / (time.perf_counter() - t0) / 1e9
        t = torch.tensor([rate, -rate], dtype=torch.float64, device=dev)
        if world > 1:
            tsum = t.clone()
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[name + "_sum_gbs"], out[name + "_min_gbs"] = float(tsum[0].item()), -float(t[1].item())
        else:
            out[name + "_sum_gbs"] = out[name + "_min_gbs"] = rate
    del h_in, h_out, d_in, d_out
    return out


def geometry_leg(fpv, synth, torch, dev, local, name, peak, steps=10):
    """Device-resident encode + decode of another BASELINE geometry, measured like the headline (CUDA events around
    the dominant kernel, inputs far larger than L2): roofline fractions for the `configs` entry of the N = 1 line."""
    W, H, bits, shift, desc = WORKLOADS[name]
    P = W * H
    F = 2368 if P < 4000000 else 1184      # one full wave of the decode kernel (16 / 8 frames per SM)
    frames = torch.empty((F, P), dtype=torch.uint16, device=dev)
    for c in range(0, F, 64):
        m = min(64, F - c)
        frames[c:c + m] = synth.plasma_frames_torch(m, W, H, bits=bits, seed=1, first=c, device=dev).reshape(m, P)
    hi = torch.empty((F, P), dtype=torch.uint8, device=dev)
    lo = torch.empty((F, P), dtype=torch.uint8, device=dev)
    pv = torch.empty((F, P // 16), dtype=torch.uint8, device=dev)
    fl = torch.empty(F, dtype=torch.uint8, device=dev)
    out = torch.empty((F, P), dtype=torch.int16, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    ctx = fpv.Context(W, H, shift, False, max_batch=F, device=local)
    ctx.set_delta_raw_device(frames[0].data_ptr(), sp)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = {"workload": desc, "xsize": W, "ysize": H, "bits": bits, "shift": shift, "frames": F}

    def enc():
        ctx.encode_device(frames.data_ptr(), F, fl.data_ptr(), hi.data_ptr(), lo.data_ptr(), pv.data_ptr(), stream=sp)

    def dec():
        ctx.decode_device(hi.data_ptr(), lo.data_ptr(), fl.data_ptr(), F, out.data_ptr(), options=fpv.DEC_UNEXTRACT, stream=sp)

    t_a = time.perf_counter()
    for what, fn, bpp in (("encode", enc, ENC_BYTES_PER_PX), ("decode", dec, DEC_BYTES_PER_PX)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ctx.enable_kernel_timing(True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        kms, kcnt = ctx.read_kernel_timing()
        ctx.enable_kernel_timing(False)
        ach = bpp * F * P / (kms / max(kcnt, 1) * 1e-3) / 1e9
        res[what] = {"value": F * P * 2 / (ms * 1e-3) / 1e9, "unit": "GB/s", "frames_per_s": F / (ms * 1e-3), "ms_per_step": ms,
                     "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                  "step_frac": bpp * F * P / (ms * 1e-3) / 1e9 / peak}}
    res["decode"]["round_trip_exact"] = bool(torch.equal(out.view(torch.uint16), frames))
    window = (t_a, time.perf_counter())
    del frames, hi, lo, pv, fl, out, ctx
    torch.cuda.empty_cache()
    return res, window


def ingest_leg(fpv_host, synth, ncpu, local, seconds):
    """BASELINE configs[4]: 16-bit 1024x1024 frames ARRIVING at a camera rate (the reference states none, so
    the offered load is swept) into fpvc::Encoder; a frame that waits longer than the camera ring holds it is dropped.
    Reports, per entropy stage, every sweep point and the highest offered rate without a drop."""
    W, H, bits, shift, _ = WORKLOADS["c1"]
    pool = synth.plasma_frames(64, W, H, bits=bits, seed=1).reshape(64, -1)
    out = {"geometry": "16-bit 1024x1024 (BASELINE configs[4])", "ring_frames": 64, "seconds_per_point": seconds,
           "latency": "frame arrival (camera clock) -> its compressed bytes reach the Encoder callback",
           "drop_rule": "a frame whose turn comes more than ring_frames / fps after its arrival is dropped, never waited for"}
    for mode, ge, batch, rates in (("host_brotli", False, 8, (500, 1000, 1500, 2000, 3000)),
                                   ("gpu_entropy", True, 16, (2000, 5000, 8000, 10000, 15000, 20000))):
        pts, best = [], None
        for fps in rates:
            r = fpv_host.ingest(pool, W, H, fps, seconds, shift=shift, threads=ncpu, batch=batch, device=local, gpu_entropy=ge,
                                ring_frames=64)
            pts.append({k: r[k] for k in ("offered_fps", "offered", "encoded", "dropped", "p50_ms", "p99_ms", "max_ms", "achieved_fps")})
            if r["dropped"] == 0:
                best = r
        out[mode] = {"batch": batch, "threads": ncpu, "sweep": pts,
                     "max_zero_drop_fps": best["offered_fps"] if best else None,
                     "max_zero_drop_raw_gbs": best["offered_fps"] * W * H * 2 / 1e9 if best else None,
                     "p50_ms_at_max": best["p50_ms"] if best else None, "p99_ms_at_max": best["p99_ms"] if best else None}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: c2 on one GPU, c3 (the 10 000-frame sequence, sharded) on several")
    ap.add_argument("--sequence-frames", type=int, default=10000, help="frames of the sharded sequence (N > 1)")
    ap.add_argument("--no-ingest", action="store_true", help="skip the paced real-time ingest leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[0] / configs[2] geometry legs (N = 1)")
    ap.add_argument("--ingest-seconds", type=float, default=1.5)
    ap.add_argument("--frames", type=int, default=2368,
                    help="frames per GPU per step (device-resident); 2368 = 148 SMs x 8 decode warps x 2 frames")
    ap.add_argument("--decode-frames", type=int, default=0, help="frames of the device-resident decode leg (0: one wave)")
    ap.add_argument("--e2e-frames", type=int, default=512, help="frames per step of the host-buffer (e2e) legs")
    ap.add_argument("--e2e-batch", type=int, default=64)
    ap.add_argument("--e2e-slots", type=int, default=3, help="overlapped submit slots of the host-buffer legs (<= 4)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    ap.add_argument("--no-stream", action="store_true", help="skip the whole-codec (brotli) leg")
    ap.add_argument("--stream-frames", type=int, default=1024)
    ap.add_argument("--no-entropy", action="store_true", help="skip the device-resident GPU entropy coder leg")
    ap.add_argument("--entropy-frames", type=int, default=256)
    args = ap.parse_args()
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    explicit_workload = args.workload is not None
    if args.workload is None:
        args.workload = "c2" if max(world_env, args.gpus) == 1 else "c3"
    W, H, bits, shift, desc = WORKLOADS[args.workload]
    P = W * H

    if args.impl == "reference":
        run_reference_arm(args, W, H, bits, shift, desc)
        return

    import torch
    import torch.distributed as dist

    import fusion_power_video_b200 as fpv
    from fusion_power_video_b200 import synth
    from fusion_power_video_b200.binding import PinnedArray

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if fpv.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from fusion_power_video_b200 import sharding

    # control-plane collectives (IPC handle, shard streams) go over gloo: no data-path collective exists
    ctl = dist.new_group(backend="gloo") if world > 1 else None
    cpus = pin_rank_cpus(local, rank, world)
    sharded = world > 1 and args.workload == "c3"
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    multi = None

    def gen(first, n):
        return synth.plasma_frames_torch(n, W, H, bits=bits, seed=1, first=first, device=dev).reshape(n, P)

    if sharded:
        # ONE sequence of --sequence-frames frames, cut into contiguous ranges; this rank's shard stays resident
        seq_total = args.sequence_frames
        f0, f1 = sharding.frame_range(seq_total, world, rank)
        F = f1 - f0
        free_b, _ = torch.cuda.mem_get_info()
        per_frame = P * 2 + 2 * P + P // 16 + 1
        fit = int(0.80 * free_b // per_frame)
        if F > fit:
            raise SystemExit(f"shard of {F} frames ({F * per_frame / 1e9:.0f} GB in + out) does not fit {free_b / 1e9:.0f} GB of HBM: "
                             "use more GPUs or a shorter --sequence-frames")
        frames = torch.empty((F, P), dtype=torch.uint16, device=dev)
        for c in range(0, F, 64):
            m = min(64, F - c)
            frames[c:c + m] = gen(f0 + c, m)
        delta = gen(0, 1).reshape(P)
        ctx = fpv.Context(W, H, shift, False, max_batch=min(F, 8192), device=local)
        # the one-frame halo: rank 0 splits the delta frame, every other rank copies the planes device to device
        t_ipc = time.perf_counter()
        if rank == 0:
            ctx.set_delta_raw_device(delta.data_ptr(), sp)
            torch.cuda.synchronize()
            box = [ctx.delta_ipc_export()]
        else:
            box = [None]
        dist.broadcast_object_list(box, src=0, group=ctl)
        if rank != 0:
            ctx.delta_ipc_import(box[0])
        dist.barrier(group=ctl)          # rank 0 keeps the exported image alive until everyone has copied it
        t_ipc = time.perf_counter() - t_ipc
        # proof that the halo arrived: frame 0 encoded against it is all zeros on every rank (frame == delta frame)
        z = [torch.empty((1, P), dtype=torch.uint8, device=dev) for _ in range(2)]
        zp = torch.empty((1, P // 16), dtype=torch.uint8, device=dev)
        zf = torch.empty(1, dtype=torch.uint8, device=dev)
        ctx.encode_device(delta.data_ptr(), 1, zf.data_ptr(), z[0].data_ptr(), z[1].data_ptr(), zp.data_ptr(), stream=sp)
        torch.cuda.synchronize()
        halo_ok = bool(int(zf.item()) & 1) and int(z[0].max().item()) == 0 and int(z[1].max().item()) == 0
        oks = [None] * world
        dist.all_gather_object(oks, halo_ok, group=ctl)
        multi = {"sequence_frames": seq_total, "frame_range_rank0": [f0, f1], "frames_per_gpu": F,
                 "delta_halo": {"how": "fpv_delta_ipc_export on rank 0 -> 64-byte handle over gloo -> fpv_delta_ipc_import "
                                       "(cudaIpcOpenMemHandle + device-to-device copy) on every other rank",
                                "bytes_per_gpu": P * 2, "seconds_incl_handshake": t_ipc,
                                "frame0_encodes_to_zero_on_every_rank": all(oks)}}
        del z, zp, zf
    else:
        F = args.frames
        seq_total = world * F
        # rank r owns frames [r*F, (r+1)*F) of the sequence; the delta frame is frame 0 (encode.cc:87-90)
        frames = torch.empty((F, P), dtype=torch.uint16, device=dev)
        for c in range(0, F, 64):
            m = min(64, F - c)
            frames[c:c + m] = gen(rank * F + c, m)
        delta = gen(0, 1).reshape(P)
        ctx = fpv.Context(W, H, shift, False, max_batch=F, device=local)
        ctx.set_delta_raw_device(delta.data_ptr(), sp)
    d_high = torch.empty((F, P), dtype=torch.uint8, device=dev)
    d_low = torch.empty((F, P), dtype=torch.uint8, device=dev)
    d_prev = torch.empty((F, P // 16), dtype=torch.uint8, device=dev)
    d_flags = torch.empty(F, dtype=torch.uint8, device=dev)

    def step():
        ctx.encode_device(frames.data_ptr(), F, d_flags.data_ptr(), d_high.data_ptr(), d_low.data_ptr(),
                          d_prev.data_ptr(), stream=sp)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    windows = []

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ctx.enable_kernel_timing(True)
    l0 = ctx.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_a = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    t_b = time.perf_counter()
    windows.append((t_a, t_b))
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.kernel_launches - l0
    kms, kcnt = ctx.read_kernel_timing()
    ctx.enable_kernel_timing(False)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = seq_total * P * 2 / (ms_step * 1e-3) / 1e9
    flags_host = d_flags.cpu().numpy()

    peak, peak_src = measured_peak()
    # one step = ceil(F / max_batch) launches of the fused kernel; their summed time covers F frames per step
    k_step_ms = kms / args.steps
    k_avg_ms = kms / max(kcnt, 1)
    achieved = ENC_BYTES_PER_PX * F * P / (k_step_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_encode_fast", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic("encode", ENC_BYTES_PER_PX * F * P),
                "traffic_source": "profiles/traffic.json (ncu --set full dram__bytes_read.sum + dram__bytes_write.sum, scaled per byte)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ENC_BYTES_PER_PX * F * P * args.steps / max(kcnt, 1), "kernel_ms": k_avg_ms,
                "launches_per_step": kcnt / args.steps, "kernel_ms_per_step": k_step_ms,
                "kernel_share_of_step": k_step_ms / ms_step if world == 1 else None}

    # ---- decode (inverse transform) on the planes just produced: extra, device-resident -------------
    decode = None
    Fd = min(F, args.decode_frames or (2368 if P < 4000000 else 1184))     # decode legs: one full wave of the decode kernel
    if not args.no_decode:
        d_out = torch.empty((Fd, P), dtype=torch.int16, device=dev)
        dsteps = max(3, min(args.steps, 20))
        for _ in range(2):
            ctx.decode_device(d_high.data_ptr(), d_low.data_ptr(), d_flags.data_ptr(), Fd, d_out.data_ptr(),
                              options=fpv.DEC_UNEXTRACT, stream=sp)
        torch.cuda.synchronize()
        ok = bool(torch.equal(d_out.view(torch.uint16), frames[:Fd]))
        ctx.enable_kernel_timing(True)
        t_a = time.perf_counter()
        e0.record(stream)
        for _ in range(dsteps):
            ctx.decode_device(d_high.data_ptr(), d_low.data_ptr(), d_flags.data_ptr(), Fd, d_out.data_ptr(),
                              options=fpv.DEC_UNEXTRACT, stream=sp)
        e1.record(stream)
        torch.cuda.synchronize()
        windows.append((t_a, time.perf_counter()))
        dms = e0.elapsed_time(e1) / dsteps
        dk, dc = ctx.read_kernel_timing()
        ctx.enable_kernel_timing(False)
        td = torch.tensor([dms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        dms = float(td.item())
        dach = DEC_BYTES_PER_PX * Fd * P / (dk / max(dc, 1) * 1e-3) / 1e9
        decode = {"metric": "decode_transform_raw_pixel_throughput", "value": world * Fd * P * 2 / (dms * 1e-3) / 1e9,
                  "unit": "GB/s", "frames_per_s": world * Fd / (dms * 1e-3), "frames_per_gpu": Fd, "ms_per_step": dms, "steps": dsteps,
                  "round_trip_exact": ok,
                  "roofline": {"bound": "hbm (chain-latency limited, see DESIGN.md)",
                               "kernel": ("k_decode_fused" if (W % 16 == 0 and 64 <= W <= 1280) else
                                          "k_decode_fused (split mode)" if (W % 32 == 0 and W <= 2560) else "k_decode_simd"),
                               "achieved": dach, "peak": peak, "unit": "GB/s", "frac": dach / peak,
                               "traffic": ncu_traffic("decode", DEC_BYTES_PER_PX * Fd * P)}}
        del d_out

    # ---- GPU entropy coder on the planes just produced: extra, device-resident ----------------------
    entropy = None
    if not args.no_entropy:
        Fn = min(F, args.entropy_frames)
        cap = ctx.stream_bound(Fn)
        d_coded = torch.empty(cap, dtype=torch.uint8, device=dev)
        d_off = torch.empty(Fn + 1, dtype=torch.int64, device=dev)
        ectx2 = fpv.Context(W, H, shift, False, max_batch=Fn, device=local)

        def ent():
            ectx2.entropy_device(d_flags.data_ptr(), d_high.data_ptr(), d_low.data_ptr(), d_prev.data_ptr(), Fn,
                                 d_coded.data_ptr(), cap, d_off.data_ptr(), stream=sp)

        for _ in range(2):
            ent()
        torch.cuda.synchronize()
        nsteps = max(3, min(args.steps, 10))
        t_a = time.perf_counter()
        e0.record(stream)
        for _ in range(nsteps):
            ent()
        e1.record(stream)
        torch.cuda.synchronize()
        windows.append((t_a, time.perf_counter()))
        ems = e0.elapsed_time(e1) / nsteps
        coded = int(d_off[Fn].item())
        entropy = {"metric": "entropy_coder_plane_throughput", "value": Fn * (2 * P + P // 16) / (ems * 1e-3) / 1e9,
                   "unit": "GB/s of plane bytes", "raw_pixel_gbs": Fn * P * 2 / (ems * 1e-3) / 1e9,
                   "frames_per_s": world * Fn / (ems * 1e-3), "ms_per_step": ems, "frames": Fn, "coded_bytes": coded,
                   "bpp": coded * 8.0 / (Fn * P),
                   "what": "fpv_entropy_device: per 64 KiB chunk histogram, Huffman code, bit packing, then layout + gather "
                           "into container chunks (valid RFC 7932 streams)"}
        del d_coded, ectx2

    # ---- e2e: host buffers through the C ABI, copies inside the timed region -----------------------
    e2e, e2e_bufs = None, ()
    if not args.no_e2e:
        big = P >= 4000000                      # 2048x2048: 8.4 MB per frame, keep the pinned staging per rank modest
        Fe, B = min(args.e2e_frames if not big else 128, F), (args.e2e_batch if not big else 16)
        Fe = (Fe // B) * B or B
        ectx = fpv.Context(W, H, shift, False, max_batch=B, device=local)
        ectx.set_delta_raw_device(delta.data_ptr(), sp)
        torch.cuda.synchronize()
        hin = PinnedArray((Fe, P), np.uint16)
        hh = PinnedArray((Fe, P), np.uint8)
        hl = PinnedArray((Fe, P), np.uint8)
        hp = PinnedArray((Fe, P // 16), np.uint8)
        hf = PinnedArray((Fe,), np.uint8)
        hin.array[:] = frames[:Fe].cpu().numpy()
        e2e_bufs = (hin, hh, hl, hp, hf)

        NS = args.e2e_slots

        def e2e_pass():
            nb = Fe // B
            for b in range(nb):
                slot = b % NS
                ectx.wait(slot)
                o = b * B
                ectx.encode_submit(slot, hin.array[o:o + B], B, hf.array[o:o + B], hh.array[o:o + B], hl.array[o:o + B],
                                   hp.array[o:o + B])
            for sl in range(NS):
                ectx.wait(sl)

        esteps = max(3, min(args.steps, 20))
        for _ in range(2):
            e2e_pass()
        if world > 1:
            dist.barrier()
        t_a = time.perf_counter()
        for _ in range(esteps):
            e2e_pass()
        t_b = time.perf_counter()
        windows.append((t_a, t_b))
        te = torch.tensor([(t_b - t_a) / esteps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
        e2e_ok = bool(np.array_equal(hh.array, d_high[:Fe].cpu().numpy()) and np.array_equal(hf.array, flags_host[:Fe]))
        e2e = {"value": world * Fe * P * 2 / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": Fe * P * 2,
               "d2h_bytes_per_step": Fe * (2 * P + P // 16 + 1), "frames_per_step": Fe, "batch": B,
               "frames_per_s": world * Fe / e2e_s, "matches_device_path": e2e_ok,
               "slots": NS,
               "what": "fpv_encode_submit/fpv_wait on pinned host buffers, slots overlapped (no brotli)"}
        # the box's own copy ceiling with every rank copying at once, measured in the same run
        pc = pcie_probe(torch, dist, world, dev)
        e2e["pcie_ceiling"] = pc
        e2e["frac_of_pcie_ceiling"] = e2e["value"] / pc["bidir_each_sum_gbs"]

    # ---- e2e with the entropy stage on the GPU: raw frames in, container chunks (coded streams) out ----------------
    e2e_stream = None
    if not args.no_e2e and not args.no_entropy:
        capb = ectx.stream_bound(B)
        hc = [PinnedArray((capb,), np.uint8) for _ in range(NS)]
        hoff = [PinnedArray((B + 1,), np.uint64) for _ in range(NS)]
        coded = [0]

        def st_pass():
            nb = Fe // B
            tot = 0
            for b in range(nb):
                slot = b % NS
                ectx.wait(slot)
                if b >= NS:
                    tot += int(hoff[slot].array[B])
                o = b * B
                ectx.encode_stream_submit(slot, hin.array[o:o + B], B, hf.array[o:o + B], hoff[slot].array, hc[slot].array, capb)
            for sl in range(NS):
                ectx.wait(sl)
            for b in range(max(0, nb - NS), nb):
                tot += int(hoff[b % NS].array[B])
            coded[0] = tot

        for _ in range(2):
            st_pass()
        if world > 1:
            dist.barrier()
        t_a = time.perf_counter()
        for _ in range(esteps):
            st_pass()
        t_b = time.perf_counter()
        windows.append((t_a, t_b))
        ts2 = torch.tensor([(t_b - t_a) / esteps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts2, op=dist.ReduceOp.MAX)
        ssec = float(ts2.item())
        e2e_stream = {"value": world * Fe * P * 2 / ssec / 1e9, "unit": "GB/s", "frames_per_s": world * Fe / ssec,
                      "h2d_bytes_per_step": Fe * P * 2, "d2h_bytes_per_step": int(coded[0]) + Fe * (1 + 8), "frames_per_step": Fe,
                      "batch": B, "slots": NS, "bpp": coded[0] * 8.0 / (Fe * P),
                      "what": "fpv_encode_stream_submit/fpv_wait on pinned host buffers: transform + GPU entropy coding + "
                              "container framing; only the coded bytes come back (about half the plane bytes)"}
        for a in hc + hoff:
            a.free()

    # ---- decode e2e: planes in pinned host memory -> fpv_decode_submit / fpv_wait -> raw frames in host memory ----
    decode_e2e = None
    if not args.no_e2e and not args.no_decode:
        ho = PinnedArray((Fe, P), np.uint16)

        def dec_pass():
            nb = Fe // B
            for b in range(nb):
                slot = b % NS
                ectx.wait(slot)
                o = b * B
                ectx.decode_submit(slot, hh.array[o:o + B], hl.array[o:o + B], hf.array[o:o + B], B, ho.array[o:o + B],
                                   options=fpv.DEC_UNEXTRACT)
            for sl in range(NS):
                ectx.wait(sl)

        for _ in range(2):
            dec_pass()
        if world > 1:
            dist.barrier()
        t_a = time.perf_counter()
        for _ in range(esteps):
            dec_pass()
        t_b = time.perf_counter()
        windows.append((t_a, t_b))
        td2 = torch.tensor([(t_b - t_a) / esteps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(td2, op=dist.ReduceOp.MAX)
        dsec = float(td2.item())
        decode_e2e = {"value": world * Fe * P * 2 / dsec / 1e9, "unit": "GB/s", "frames_per_s": world * Fe / dsec,
                      "h2d_bytes_per_step": Fe * (2 * P + 1), "d2h_bytes_per_step": Fe * P * 2, "frames_per_step": Fe, "batch": B,
                      "round_trip_exact": bool(np.array_equal(ho.array, hin.array)),
                      "what": "fpv_decode_submit/fpv_wait on pinned host buffers (planes in, raw file words out), slots overlapped (a batch's kernel alone takes 1 ms: the chain)"}
        ho.free()

    # ---- whole codec: fpvc::Encoder (GPU transform + host brotli + framing) on host frames, every rank at once ------
    stream_leg = None
    if not args.no_e2e and not args.no_stream:
        from fusion_power_video_b200 import host as fpv_host

        ncpu = len(cpus)                      # this rank's share of the host cores (all of them at N = 1)
        big = P >= 4000000
        ns = args.stream_frames if not big else 96
        if world > 1:
            ns = max(32, ns // world)
        fr = np.ascontiguousarray(np.tile(hin.array, ((ns + Fe - 1) // Fe, 1))[:ns])   # the stream repeats the e2e frames

        def timed_encode(n_rep, frames_in=None, **kw):
            """best-of-n_rep seconds of Init + CompressFrame x ns + Finish, all ranks started together; max over ranks"""
            best, size = None, 0
            for _ in range(n_rep):
                if world > 1:
                    dist.barrier(group=ctl)
                t, size = fpv_host.time_encode(fr if frames_in is None else frames_in, W, H, shift, False, threads=ncpu,
                                               device=local, **kw)
                tt = [None] * world
                if world > 1:
                    dist.all_gather_object(tt, t, group=ctl)
                    t = max(tt)
                best = t if best is None else min(best, t)
            return best, size

        fpv_host.time_encode(fr[:2], W, H, shift, False, threads=ncpu, batch=8, device=local)   # warm-up: context, pinned pools
        t_a = time.perf_counter()
        best, size = timed_encode(3 if world == 1 else 2, batch=8)
        windows.append((t_a, time.perf_counter()))
        stream_leg = {"what": "fpvc::Encoder Init + CompressFrame x n + Finish (benchmark.cc:153-180 window): pinned H2D, "
                              "GPU transform, D2H, brotli q1 on host threads, framing; one Encoder per rank / GPU, all at once",
                      "value": world * ns * P * 2 / best / 1e9, "unit": "GB/s", "frames_per_s": world * ns / best,
                      "mp_per_s": world * ns * P / best / 1e6, "frames": world * ns, "host_threads": ncpu * world,
                      "stream_bytes_rank0": int(size), "bpp": size * 8.0 / (ns * P),
                      "bound": "host brotli (about 70 MP/s per core) -- the transform is off the critical path"}

        # the same codec with the entropy stage on the GPU (brotli-compatible streams, no host brotli)
        fpv_host.time_encode(fr[:32], W, H, shift, False, threads=ncpu, batch=32 if not big else 8, gpu_entropy=True, device=local)
        t_a = time.perf_counter()
        bestg, sizeg = timed_encode(5 if world == 1 else 2, batch=32 if not big else 8, gpu_entropy=True)
        windows.append((t_a, time.perf_counter()))
        stream_leg["gpu_entropy"] = {
            "what": "same Encoder with GpuOptions::gpu_entropy: transform + chunk-parallel Huffman coding (valid RFC 7932 "
                    "streams the reference decoder reads) + framing on the GPU; D2H carries only the coded bytes",
            "value": world * ns * P * 2 / bestg / 1e9, "unit": "GB/s", "frames_per_s": world * ns / bestg,
            "mp_per_s": world * ns * P / bestg / 1e6, "stream_bytes_rank0": int(sizeg), "bpp": sizeg * 8.0 / (ns * P),
            "batch": 32 if not big else 8, "bound": "the Encoder's copy of every frame into pinned memory, then PCIe"}
        # ... and with the caller's frames already in page-locked memory (camera DMA buffers): the Encoder uploads them
        # from where they are, no host copy (the reference's contract: img stays valid until its callback)
        if world == 1 or not big:
            fpin = PinnedArray((ns, P), np.uint16)
            fpin.array[:] = fr
            t_a = time.perf_counter()
            bestp, sizep = timed_encode(5 if world == 1 else 2, frames_in=fpin.array, batch=32 if not big else 8, gpu_entropy=True)
            windows.append((t_a, time.perf_counter()))
            stream_leg["gpu_entropy"]["pinned_input"] = {
                "value": world * ns * P * 2 / bestp / 1e9, "unit": "GB/s", "frames_per_s": world * ns / bestp,
                "stream_identical": bool(sizep == sizeg), "bound": "PCIe: 2 B/px in, about 1 B/px out",
                "what": "frames handed to CompressFrame in page-locked memory: zero host copies"}
            del fpin

        # decode side of the codec: StreamingDecoder (host brotli decode on all cores + GPU inverse transform +
        # UnextractFrame on the GPU) on the stream just described, both entropy variants (rank 0, N = 1)
        if world == 1:
            nd = min(ns, 1024 if not big else 32)
            dec = {}
            for name, ge in (("brotli_stream", False), ("gpu_entropy_stream", True)):
                st = fpv_host.encode_stream(fr[:nd], W, H, shift, False, threads=ncpu, batch=32, gpu_entropy=ge)
                db = 64 if not big else 16       # frames per GPU call of the decoder (GpuOptions::batch)
                out = fpv_host.decode_stream(st, nd, W, H, block=0, batch=db, raw_shift=shift)
                okd = bool(np.array_equal(out, fr[:nd]))
                del out
                t_a = time.perf_counter()
                bestd = bestc = None
                for _ in range(2):
                    # the decoder alone (the callback counts frames) and with a consumer that copies every frame out
                    cnt, sec, first = fpv_host.decode_stream(st, nd, W, H, block=0, batch=db, raw_shift=shift, return_time="both",
                                                             keep=False)
                    okd = okd and cnt.shape[0] == nd
                    if bestd is None or sec < bestd:
                        bestd, steady = sec, (nd - db) / max(sec - first, 1e-9)
                    _, sec = fpv_host.decode_stream(st, nd, W, H, block=0, batch=db, raw_shift=shift, return_time=True)
                    bestc = sec if bestc is None else min(bestc, sec)
                windows.append((t_a, time.perf_counter()))
                dec[name] = {"value": nd * P * 2 / bestd / 1e9, "unit": "GB/s", "frames_per_s": nd / bestd, "frames": nd,
                             "with_consumer_copy": nd * P * 2 / bestc / 1e9, "batch": db,
                             "steady_frames_per_s": steady, "steady_gbs": steady * P * 2 / 1e9,
                             "what": "value: the whole call, decoder construction (GPU context, pinned staging) included; steady: "
                                     "frames after the first batch / time after the first frame came out",
                             "round_trip_exact": okd, "stream_bytes": len(st)}
            stream_leg["decode"] = dec

    # ---- merged-stream check (N > 1): shards of a subsample of the sequence, encoded per rank, merge to the stream
    #      one GPU writes for the same frames ------------------------------------------------------------------------
    if sharded and not args.no_e2e:
        import hashlib
        from fusion_power_video_b200 import host as fpv_host

        n_sub = 4 * world
        idx = [k * (seq_total // n_sub) for k in range(n_sub)]            # spread over the whole sequence; idx[0] = the delta frame
        sa, sb = sharding.frame_range(n_sub, world, rank)
        mine = torch.cat([gen(idx[j], 1) for j in range(sa, sb)]).cpu().numpy()
        delta_np = delta.cpu().numpy()
        local_stream = fpv_host.encode_stream(mine, W, H, shift, False, threads=len(cpus), batch=4, delta=delta_np, device=local)
        merged = sharding.gather_stream(local_stream, group=ctl)
        if rank == 0:
            every = torch.cat([gen(i, 1) for i in idx]).cpu().numpy()
            single = fpv_host.encode_stream(every, W, H, shift, False, threads=len(cpus), batch=4, delta=delta_np, device=local)
            back = fpv_host.decode_stream(merged, n_sub, W, H, block=0, batch=4, raw_shift=shift)
            multi["merged_stream"] = {
                "frames": n_sub, "sequence_indices": idx, "bytes": len(merged),
                "sha256_merged": hashlib.sha256(merged).hexdigest(), "sha256_single_gpu": hashlib.sha256(single).hexdigest(),
                "matches_single_gpu_stream": merged == single, "round_trip_exact": bool(np.array_equal(back, every)),
                "what": "every rank encodes its range of the subsample with fpvc::Encoder (host brotli), rank 0 merges the "
                        "shard streams (sharding.merge_shards: chunks in rank order, footer offsets = prefix sum) and compares "
                        "with the stream its own single GPU writes for all of the subsample"}

    # ---- the other BASELINE geometries, same measurement (N = 1 line only) -------------------------------------------
    configs = None
    if world == 1 and not args.no_configs and not explicit_workload:
        peak_c, _ = measured_peak()
        configs = {}
        for nm in ("c1", "c3"):
            configs[nm], wnd = geometry_leg(fpv, synth, torch, dev, local, nm, peak_c)
            windows.append(wnd)

    # ---- BASELINE configs[4]: paced real-time ingest (N = 1 line only) ------------------------------------------------
    ingest = None
    if world == 1 and not args.no_ingest and not args.no_e2e and not explicit_workload:
        from fusion_power_video_b200 import host as fpv_host

        ingest = ingest_leg(fpv_host, synth, len(cpus), local, args.ingest_seconds)

    clocks = sampler.stop(windows) if rank == 0 else None

    # ---- CPU baseline: the reference's own code on this box's host cores (rank 0, N == 1 only) -----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        ncpu = os.cpu_count() or 1
        ns = min(F, 4 * ncpu)
        fr = frames[:ns].cpu().numpy()
        dl = delta.cpu().numpy()
        rate, kind, cores, reps = cpu_reference_rate(fr, W, H, shift, dl, args.cpu_seconds)
        cpu_baseline = {"value": rate, "unit": "GB/s", "cores": cores, "kind": kind,
                        "sample": f"{ns} frames {W}x{H}, Frame ctor + Frame::Predict, best of {reps} passes"}
        if stream_leg is not None:
            from oracle_binding import Ref, ref_available

            if ref_available():
                nsr = min(stream_leg["frames"], 16 * ncpu)
                frs = np.ascontiguousarray(np.tile(hin.array, ((nsr + Fe - 1) // Fe, 1))[:nsr])
                t, size = Ref().time_encode(frs, W, H, shift, 0, frs[0], ncpu)
                if "decode" in stream_leg:
                    ndr = min(64, nsr)
                    stb = Ref().encode_stream(frs[:ndr], W, H, shift, 0, frs[0], ncpu)
                    t0 = time.perf_counter()
                    nrd, _, _, _ = Ref().decode_stream(stb, ndr, W, H)
                    tdr = time.perf_counter() - t0
                    stream_leg["decode"]["reference_cpu"] = {
                        "value": ndr * P * 2 / tdr / 1e9, "unit": "GB/s", "frames_per_s": ndr / tdr, "frames": int(nrd),
                        "what": "unmodified reference StreamingDecoder (single-threaded by design) on the reference's stream"}
                stream_leg["reference_cpu"] = {"value": nsr * P * 2 / t / 1e9, "unit": "GB/s", "frames_per_s": nsr / t,
                                               "mp_per_s": nsr * P / t / 1e6, "frames": nsr, "threads": ncpu,
                                               "what": "unmodified reference Encoder (transform + brotli) on the same host"}

    for a in e2e_bufs:
        a.free()
    if rank == 0:
        uniq, cnt = np.unique(flags_host, return_counts=True)
        line = {
            "metric": "encode_transform_raw_pixel_throughput", "value": value, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "frames_per_s": seq_total / (ms_step * 1e-3),
            "config": {"workload": desc, "xsize": W, "ysize": H, "bits": bits, "shift": shift,
                       "frames_per_gpu_per_step": F, "frames_per_step": seq_total,
                       "parallelism": (f"one {seq_total}-frame sequence cut into {world} contiguous frame ranges, delta frame by peer copy (CUDA IPC), no collective"
                                       if sharded else f"frame-range x{world}, no collective"),
                       "l2": f"inputs larger than L2: {F * P * 2 / 1e6:.0f} MB raw + {F * P * 2.0625 / 1e6:.0f} MB out per step vs 126 MB L2",
                       "flags_histogram": {int(u): int(c) for u, c in zip(uniq, cnt)}},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "decode": decode, "e2e_stream": e2e_stream, "decode_e2e": decode_e2e, "entropy": entropy, "stream": stream_leg,
            "multi_gpu": multi, "configs": configs, "ingest": ingest,
            "host": {"cpus_per_rank": len(cpus), "cpu_count": os.cpu_count()},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
