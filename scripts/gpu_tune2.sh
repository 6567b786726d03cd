#!/bin/bash
# Sweep env knobs given as "NAME=VALUE,NAME=VALUE" groups with the kernel-only encode bench.
mkdir -p gpurun_out
python -m pytest tests/test_encode_gpu.py -m gpu -x -q 2>&1 | tail -3
for cfg in "$@"; do
  env $(echo $cfg | tr ',' ' ') python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --no-decode ${WORKLOAD:+--workload $WORKLOAD} > gpurun_out/tune.json 2> gpurun_out/tune.err
  python - "$cfg" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/tune.json').read().strip().splitlines()[-1])
    r=d["roofline"]
    print(sys.argv[1], "step ms", round(d["ms_per_step"],4), "kernel ms", round(r["kernel_ms"],4), "achieved", round(r["achieved"],1), "frac", round(r["frac"],3))
except Exception as e:
    print(sys.argv[1], "failed", e, open('gpurun_out/tune.err').read()[-500:])
PY
done
