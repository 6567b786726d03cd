#!/bin/bash
# Throughput vs frames in flight.  Args: frame counts
for F in "$@"; do
  python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --frames $F > gpurun_out/frames.json 2> gpurun_out/frames.err
  python - $F <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/frames.json').read().strip().splitlines()[-1])
    r=d["roofline"]; de=d["decode"]
    print("F", sys.argv[1], "enc kernel frac", round(r["frac"],3), "enc step ms", round(d["ms_per_step"],3), "| dec ms", round(de["ms_per_step"],3), "dec frac", round(de["roofline"]["frac"],3), "exact", de["round_trip_exact"])
except Exception as e:
    print(sys.argv[1], "failed", e, open('gpurun_out/frames.err').read()[-800:])
PY
done
