#!/bin/bash
# Round-2 one-GPU visit: parity suite (incl. the reference's programs on the library), the new N = 1 bench line, reference arm.
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
( time timeout -s KILL 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err ) 2>&1 | tail -3
tail -c 1500 gpurun_out/bench_r2_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r2_n1.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "frac", round(d["roofline"]["frac"], 3), "share", d["roofline"]["kernel_share_of_step"], "e2e", round(d["e2e"]["value"], 1),
          "decode frac", round(d["decode"]["roofline"]["frac"], 3), "launches", d["gpu_launches"])
    for k, v in (d.get("configs") or {}).items():
        print(k, "enc", round(v["encode"]["roofline"]["frac"], 3), "step", round(v["encode"]["roofline"]["step_frac"], 3), "dec", round(v["decode"]["roofline"]["frac"], 3), v["decode"]["round_trip_exact"])
    ing = d.get("ingest") or {}
    for m in ("host_brotli", "gpu_entropy"):
        if m in ing:
            print(m, "max zero-drop fps", ing[m]["max_zero_drop_fps"], "p50", ing[m]["p50_ms_at_max"], "p99", ing[m]["p99_ms_at_max"])
            for p in ing[m]["sweep"]:
                print("   ", p)
    print("stream", round(d["stream"]["value"], 2), "gpu_entropy", round(d["stream"]["gpu_entropy"]["value"], 2))
except Exception as e:
    print("bench parse failed", e)
PY
timeout -s KILL 200 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-300
