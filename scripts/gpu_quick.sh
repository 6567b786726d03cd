#!/bin/bash
# Quick GPU visit: parity tests + short bench (+ optional ncu of the main encode pass when $1 == prof).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
    print("value GB/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "roofline", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in("achieved","frac","kernel_ms","kernel_share_of_step")})
    print("decode", d["decode"] and {k:d["decode"][k] for k in ("value","ms_per_step","round_trip_exact")}, d["decode"] and d["decode"]["roofline"]["frac"])
    print("e2e", d["e2e"] and d["e2e"]["value"], "flags", d["config"]["flags_histogram"], "launches", d["gpu_launches"], d["clocks"])
except Exception as e:
    print("bench failed", e); print(open('gpurun_out/bench_quick.err').read()[-3000:])
PY
if [ "$1" == "prof" ]; then
  ncu --set full --clock-control none --import-source on -k regex:k_encode_fast -s 3 -c 1 -o gpurun_out/prof_encode_main -f \
    python bench.py --steps 3 --warmup 3 --frames 512 --no-e2e --no-cpu --no-decode > gpurun_out/ncu_encode.log 2>&1
  K='regex:k_(encode|decode|decide|finalize|gen|delta|cg|combine|planes)'
  ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --frames 512 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
fi
