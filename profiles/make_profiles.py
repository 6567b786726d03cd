#!/usr/bin/env python
"""Turns the ncu artefacts a `scripts/gpu_profile.sh` visit leaves in gpurun_out/ into the tracked
summaries under profiles/ (run here, where ncu is installed but no GPU is):

    python profiles/make_profiles.py r01

  profiles/<round>_launches.csv      the launch list (gpu__time_duration.sum per launch) of one bench command
  profiles/<round>_launches.md       per-kernel share of that command
  profiles/<round>_encode.md         k_encode_fast main pass: ncu --set full metrics + SASS opcode mix + top stalls
  profiles/<round>_decode.md         k_decode_pair: same
"""
import collections
import csv
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
sys.path.insert(0, os.path.join(ROOT, "profiles"))
from ncu_summary import KEYS  # noqa: E402

EXTRA = ["dram__bytes_read.sum", "dram__bytes_write.sum", "launch__occupancy_limit_warps", "sm__maximum_warps_per_active_cycle_pct"]


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"name": r[hdr.index("Kernel Name")]}
        for k in KEYS + EXTRA:
            if k in hdr:
                d[k] = (r[hdr.index(k)], units[hdr.index(k)])
        res.append(d)
    return res


def sass_mix(rep, units_per_launch, unit_name):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ops, samp = collections.Counter(), collections.Counter()
    top = []
    total = 0
    for r in rows[2:]:
        if len(r) <= ie:
            continue
        n, s = int(r[ie] or 0), int(r[isamp] or 0)
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
        op = m.group(2).split(".")[0] if m else "?"
        ops[op] += n
        samp[op] += s
        total += n
        top.append((s, n, r[ia].strip()))
    lines = [f"Warp instructions executed: {total:,} = {total / units_per_launch:.1f} per {unit_name}", "",
             f"| opcode | per {unit_name} | stall samples |", "|---|---|---|"]
    for k, v in ops.most_common(16):
        lines.append(f"| {k} | {v / units_per_launch:.1f} | {samp[k]} |")
    lines += ["", "Most-sampled instructions (warp stall sampling):", "", "| samples | executed | SASS |", "|---|---|---|"]
    for s, n, src in sorted(top, reverse=True)[:10]:
        lines.append(f"| {s} | {n:,} | `{src[:80]}` |")
    return "\n".join(lines)


def kernel_md(tag, title, rep, units_per_launch, unit_name, algo_bytes, note):
    ms = raw_metrics(rep)[0]
    dur_us = float(ms["gpu__time_duration.sum"][0]) * (1000.0 if ms["gpu__time_duration.sum"][1] == "ms" else 1.0)
    rd = float(ms["dram__bytes_read.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[ms["dram__bytes_read.sum"][1]]
    wr = float(ms["dram__bytes_write.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[ms["dram__bytes_write.sum"][1]]
    lines = [f"# {title}", "", note, "",
             f"Kernel: `{ms['name']}`", "",
             f"* duration under ncu (cold caches, serialised): {dur_us:.1f} us",
             f"* algorithmic bytes per launch: {algo_bytes / 1e9:.3f} GB; DRAM traffic (`dram__bytes_read.sum + dram__bytes_write.sum`): "
             f"{rd / 1e9:.3f} + {wr / 1e9:.3f} = {(rd + wr) / 1e9:.3f} GB ({(rd + wr) / algo_bytes:.2f}x algorithmic)",
             f"* algorithmic bandwidth under ncu: {algo_bytes / dur_us / 1e3:.0f} GB/s (the bench value is taken WITHOUT the profiler, see BENCH / DESIGN.md)",
             "", "| metric | value |", "|---|---|"]
    for k in KEYS:
        if k in ms:
            lines.append(f"| `{k}` | {ms[k][0]} {ms[k][1]} |")
    lines += ["", "## SASS mix", "", sass_mix(rep, units_per_launch, unit_name), ""]
    open(os.path.join(ROOT, "profiles", f"{tag}.md"), "w").write("\n".join(lines))
    return {"kernel": ms["name"], "dram_bytes_per_launch": rd + wr, "algorithmic_bytes_per_launch": algo_bytes,
            "dram_bytes_per_algorithmic_byte": (rd + wr) / algo_bytes, "ncu_duration_us": dur_us}


def launches_md(tag):
    src = os.path.join(OUT, "launches.csv")
    shutil.copy(src, os.path.join(ROOT, "profiles", f"{tag}_launches.csv"))
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    per = collections.OrderedDict()
    for r in rows[1:]:
        per.setdefault(r[ik].split("(")[0], []).append(float(r[iv]) / 1000)
    tot = sum(sum(v) for v in per.values())
    lines = [f"# Launch list ({tag})", "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_... -c 120 python bench.py --steps 3 --warmup 3 "
             "--no-e2e --no-cpu --no-stream` on one B200 (the bench's default workload, C2: 12-bit 1280x800, one full wave of frames per step; "
             "encode steps, then decode steps, then the entropy coder on 256 frames).  Per-launch times are cold-cache and "
             "serialised: compare shares, not absolutes.", "",
             "| kernel | launches | avg us | min us | max us | share of all listed time |", "|---|---|---|---|---|---|"]
    for k, v in per.items():
        lines.append(f"| `{k}` | {len(v)} | {sum(v) / len(v):.1f} | {min(v):.1f} | {max(v):.1f} | {100 * sum(v) / tot:.1f} % |")
    enc = {k: v for k, v in per.items() if "decode" not in k and "delta_" not in k and "entropy" not in k}
    etot = sum(sum(v) for v in enc.values())
    main = [x for k, v in enc.items() if "k_encode_fast" in k for x in v if x > 100]
    lines += ["", f"Within the encode steps, the main pass of `k_encode_fast` (the {len(main)} launches > 100 us) is "
              f"{100 * sum(main) / etot:.1f} % of the listed encode time; bench.py's event timing of the same share "
              "(`roofline.kernel_share_of_step`) is in the bench line."]
    open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w").write("\n".join(lines) + "\n")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    P = 1280 * 800
    import json
    if tag != "r01":
        return main_r02(tag)
    launches_md(tag)
    traffic = {}
    F = 1184
    traffic["encode"] = kernel_md(f"{tag}_encode", f"k_encode_fast, main pass (C2, {F} frames)", os.path.join(OUT, "prof_encode_main.ncu-rep"),
              F * P / 256, "warp-row (256 px)", F * P * 4.0625,
              "`ncu --set full --clock-control none --import-source on -k regex:k_encode_fast -s 3 -c 1 python bench.py --steps 3 "
              "--warmup 3 --no-e2e --no-cpu --no-stream --no-decode --no-entropy` (the 4th matching launch = the main pass of the 2nd step).")
    traffic["decode"] = kernel_md(f"{tag}_decode", f"k_decode_pair (C2, {F} frames)", os.path.join(OUT, "prof_decode.ncu-rep"),
              F * 800, "frame row (1280 px)", F * P * 4.0,
              "`ncu --set full --clock-control none --import-source on -k regex:k_decode_pair -s 1 -c 1 python bench.py --steps 3 "
              "--warmup 3 --no-e2e --no-cpu --no-stream --no-entropy`.")
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)


def main_r02(tag):
    """Round 2: the bench's default step is 2368 frames of C2 (one wave of k_decode_fused: 16 frames per SM); the
    captures come from scripts/gpu_profile_r2.sh.  Missing reports are skipped."""
    import json
    P, F = 1280 * 800, 2368
    cmd = "python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-stream --no-configs --no-ingest"
    if os.path.exists(os.path.join(OUT, "launches.csv")):
        launches_md(tag)
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}

    def have(name):
        return os.path.exists(os.path.join(OUT, name))

    if have("prof_encode_main.ncu-rep"):
        traffic["encode"] = kernel_md(f"{tag}_encode", f"k_encode_fast, main pass (C2, {F} frames)", os.path.join(OUT, "prof_encode_main.ncu-rep"),
              F * P / 256, "warp-row (256 px)", F * P * 4.0625,
              f"`ncu --set full --clock-control none --import-source on -k regex:k_encode_fast -s 9 -c 1 {cmd} --no-decode --no-entropy` "
              "(a main pass of a timed step; decisions, flags and preview prediction are fused into this kernel since round 2).")
    if have("prof_decode.ncu-rep"):
        traffic["decode"] = kernel_md(f"{tag}_decode", f"k_decode_fused (C2: 12-bit 1280x800, L = 40, bulk-store output, {F} frames = one wave)",
              os.path.join(OUT, "prof_decode.ncu-rep"), F * 800 / 2, "pair row (2 x 1280 px)", F * P * 4.0,
              f"`ncu --set full --clock-control none --import-source on -k regex:k_decode_fused -s 2 -c 1 {cmd} --no-entropy`.")
    if have("prof_decode_c1.ncu-rep"):
        P1 = 1024 * 1024
        traffic["decode_c1"] = kernel_md(f"{tag}_decode_c1", f"k_decode_fused (C1 geometry: 16-bit 1024x1024, L = 32, 256-bit direct stores, {F} frames)",
              os.path.join(OUT, "prof_decode_c1.ncu-rep"), F * 1024 / 2, "pair row (2 x 1024 px)", F * P1 * 4.0,
              f"`ncu ... -k regex:k_decode_fused -s 2 -c 1 {cmd} --workload c1 --no-entropy`.")
    if have("prof_decode_c3.ncu-rep"):
        F3, P3 = 1184, 2048 * 2048
        traffic["decode_c3"] = kernel_md(f"{tag}_decode_c3", f"k_decode_fused in split mode (C3: 16-bit 2048x2048, {F3} frames = one wave)",
              os.path.join(OUT, "prof_decode_c3.ncu-rep"), F3 * 2048, "frame row (2048 px)", F3 * P3 * 4.0,
              f"`ncu ... -k regex:k_decode_fused -s 2 -c 1 {cmd} --workload c3 --no-entropy`.")
    if have("prof_encode_c3.ncu-rep"):
        F3, P3 = 2368, 2048 * 2048          # the encode leg of `--workload c3` keeps the default 2368 frames per step
        traffic["encode_c3"] = kernel_md(f"{tag}_encode_c3", f"k_encode_fast, main pass (C3: 16-bit 2048x2048, {F3} frames)",
              os.path.join(OUT, "prof_encode_c3.ncu-rep"), F3 * P3 / 256, "warp-row (256 px)", F3 * P3 * 4.0625,
              f"`ncu ... -k regex:k_encode_fast -s 9 -c 1 {cmd} --workload c3 --no-decode --no-entropy`.")
    if have("prof_entropy.ncu-rep"):
        Fe = 256
        traffic["entropy"] = kernel_md(f"{tag}_entropy", f"k_entropy_chunk (C2 planes, {Fe} frames = {Fe * 33} chunks of 64 KiB)",
              os.path.join(OUT, "prof_entropy.ncu-rep"), Fe * 33, "chunk (<= 64 KiB)", Fe * P * (2 + 1 / 16) * 1.5,
              f"`ncu ... -k regex:k_entropy_chunk -s 1 -c 1 {cmd} --no-decode`.  Algorithmic bytes here = planes read once + coded "
              "bytes written to the scratch (about half the plane bytes).  Not HBM bound: the serial Huffman construction per "
              "chunk dominates.")
    if have("prof_entropy_decode.ncu-rep"):
        Fd = 64
        traffic["entropy_decode"] = kernel_md(f"{tag}_entropy_decode", f"k_entropy_decode (C2, {Fd} frames = {Fd * 32} chunks, one warp each)",
              os.path.join(OUT, "prof_entropy_decode.ncu-rep"), Fd * 32, "chunk (<= 64 KiB)", Fd * P * 2 * 1.5,
              "`ncu ... -k regex:k_entropy_decode -s 2 -c 1 python scripts/gpu_entdec.py 64`.  Algorithmic bytes = coded bytes "
              "read (about half the plane bytes) + plane bytes written.")
    json.dump(traffic, open(tpath, "w"), indent=1)


if __name__ == "__main__":
    main()
