#!/bin/bash
# Device-resident encode / decode of the other BASELINE geometries (c1: 1024x1024 16-bit, c3: 2048x2048 16-bit).
mkdir -p gpurun_out
for WL in ${WORKLOADS:-c1 c3}; do
  F=${FRAMES:-1184}
  python bench.py --workload $WL --frames $F --steps 10 --warmup 3 --no-cpu --no-e2e --no-stream --no-entropy > gpurun_out/bench_$WL.json 2> gpurun_out/bench_$WL.err
  python - $WL <<'PY'
import json,sys
WL=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/bench_{WL}.json').read().strip().splitlines()[-1])
    print(WL, "frames", d["config"]["frames_per_gpu_per_step"], "encode GB/s", round(d["value"],1), "kernel frac", round(d["roofline"]["frac"],3), "flags", d["config"]["flags_histogram"],
          "| decode GB/s", round(d["decode"]["value"],1), "frac", round(d["decode"]["roofline"]["frac"],3), d["decode"]["roofline"]["kernel"], "exact", d["decode"]["round_trip_exact"])
except Exception as e:
    print(WL, "failed", e, open(f'gpurun_out/bench_{WL}.err').read()[-800:])
PY
done
