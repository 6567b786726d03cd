#!/bin/bash
# encode ring sweep: FPV_STAGES x FPV_ROWS_PER_STAGE (x FPV_MAX_CTAS) per workload
mkdir -p gpurun_out
for WL in ${WLS:-c3 c1 c2}; do for CFG in ${CFGS:-"2 4" "3 4" "4 4" "2 2" "3 2" "4 2" "6 2"}; do
  set -- $CFG
  FPV_STAGES=$1 FPV_ROWS_PER_STAGE=$2 timeout -s KILL 120 python bench.py --steps 20 --warmup 3 --workload $WL --no-decode --no-e2e --no-cpu --no-stream --no-entropy --no-configs --no-ingest > gpurun_out/enc_tune.json 2> gpurun_out/enc_tune.err
  python - $WL $1 $2 <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/enc_tune.json').read().strip().splitlines()[-1])
    r=d["roofline"]
    print(*sys.argv[1:], "step ms", round(d["ms_per_step"],4), "kernel ms", round(r["kernel_ms"],4), "frac", round(r["frac"],4))
except Exception as e:
    print(*sys.argv[1:], "failed", e, open('gpurun_out/enc_tune.err').read()[-300:])
PY
done; done
