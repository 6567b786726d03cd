#!/usr/bin/env python
"""Host <-> device copy ceilings of the box with N GPUs busy AT THE SAME TIME (one process per GPU), with
and without NUMA pinning of each process to its GPU's node.  This is the bound of the host-buffer (`e2e`)
legs of bench.py: they move 2 B/px in and 2.06 B/px out per GPU concurrently.

    python scripts/gpu_pcie_multi.py [--mb 128] [--seconds 1.5] [--out gpurun_out/pcie_multi.json]

Prints a table (per-GPU min / aggregate GB/s for H2D alone, D2H alone, both directions at once) for
N = 1, 2, 4, 8 <= visible GPUs and the topology the numbers were taken on."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gpu_numa_cpus(index):
    """(numa node, cpu list) of GPU `index` from sysfs; (-1, None) when the platform does not say (e.g. a VM)."""
    try:
        bus = subprocess.check_output(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                                      text=True).strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return -1, None
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        out = []
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            out.extend(range(int(a), int(b or a) + 1))
        return node, out
    except Exception:
        return -1, None


def pin_to_gpu_node(index):
    node, cpus = gpu_numa_cpus(index)
    if cpus:
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
    return node


def child(rank, world, mb, seconds, pin, barrier, q):
    import torch

    node = pin_to_gpu_node(rank) if pin else -1
    torch.cuda.set_device(rank)
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out.fill_(2)
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

    res = {"rank": rank, "numa_node": node}
    for name, fn in (("h2d", h2d), ("d2h", d2h), ("bidir_each", both)):
        fn()
        torch.cuda.synchronize()
        barrier.wait()
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < seconds:
            for _ in range(4):
                fn()
            torch.cuda.synchronize()
            reps += 4
        dt = time.perf_counter() - t0
        res[name] = n * reps / dt / 1e9
        barrier.wait()
    q.put(res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=128)
    ap.add_argument("--seconds", type=float, default=1.5)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "pcie_multi.json"))
    args = ap.parse_args()
    import torch
    import torch.multiprocessing as mp

    ngpu = torch.cuda.device_count()
    topo = {}
    for name, cmd in (("nvidia_smi_topo", ["nvidia-smi", "topo", "-m"]), ("lscpu", ["lscpu"]),
                      ("numa_nodes", ["sh", "-c", "ls -d /sys/devices/system/node/node* 2>/dev/null; cat /sys/devices/system/node/node*/cpulist 2>/dev/null"]),
                      ("meminfo", ["sh", "-c", "grep -E 'MemTotal|MemFree' /proc/meminfo"])):
        try:
            topo[name] = subprocess.check_output(cmd, text=True, stderr=subprocess.STDOUT)
        except Exception as e:
            topo[name] = f"unavailable: {e}"
    topo["gpu_numa"] = {i: gpu_numa_cpus(i)[0] for i in range(ngpu)}
    topo["nproc"] = os.cpu_count()
    ctx = mp.get_context("spawn")
    rows = []
    for world in (1, 2, 4, 8):
        if world > ngpu:
            break
        for pin in (False, True):
            barrier = ctx.Barrier(world)
            q = ctx.Queue()
            ps = [ctx.Process(target=child, args=(r, world, args.mb, args.seconds, pin, barrier, q)) for r in range(world)]
            for p in ps:
                p.start()
            res = [q.get(timeout=300) for _ in ps]
            for p in ps:
                p.join()
            row = {"gpus": world, "numa_pinned": pin, "chunk_mb": args.mb}
            for k in ("h2d", "d2h", "bidir_each"):
                vals = [r[k] for r in res]
                row[k + "_min_gbs"] = min(vals)
                row[k + "_sum_gbs"] = sum(vals)
            rows.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"rows": rows, "topology": topo}, f, indent=1)
    print(topo["nvidia_smi_topo"])
    print("gpu numa nodes:", topo["gpu_numa"], "nproc:", topo["nproc"])


if __name__ == "__main__":
    main()
