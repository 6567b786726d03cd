#!/bin/bash
# Sweep of the pair kernel's tuning knobs (FPV_PAIR_TUNE="K0,G") on the decode leg of the bench.
mkdir -p gpurun_out
for T in ${TUNES:-"16,4" "16,8" "24,4" "24,8" "8,4" "12,4" "20,4"}; do for F in ${FRAMES:-1184}; do
  FPV_PAIR_TUNE=$T timeout -s KILL 120 python bench.py --steps 20 --warmup 3 --frames $F --no-cpu --no-e2e --no-stream > gpurun_out/tune.json 2> gpurun_out/tune.err
  python - "$T" "$F" <<'PY'
import json,sys
T,F=sys.argv[1:3]
try:
    d=json.loads(open('gpurun_out/tune.json').read().strip().splitlines()[-1]); dd=d["decode"]
    print("tune",T,"F",F,"ms",round(dd["ms_per_step"],4),"exact",dd["round_trip_exact"],"frac",round(dd["roofline"]["frac"],4))
except Exception as e:
    print(T,F,"failed",e); print(open('gpurun_out/tune.err').read()[-1500:])
PY
done; done
