#!/bin/bash
# decode leg at several frame counts (frames in flight per launch) for each kernel in KERNELS
mkdir -p gpurun_out
for K in ${KERNELS:-fused pair}; do for WL in ${WLS:-c2 c1 c3}; do for F in ${FRAMES:-1184 2368}; do
  if [ $WL = c3 ]; then FF=$((F/2)); else FF=$F; fi
  FPV_DECODE_KERNEL=$K timeout -s KILL 120 python bench.py --steps 10 --warmup 3 --workload $WL --frames $FF --decode-frames $FF --no-e2e --no-cpu --no-stream --no-entropy --no-configs --no-ingest > gpurun_out/dec_iter.json 2> gpurun_out/dec_iter.err
  python - $K $WL $FF <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/dec_iter.json').read().strip().splitlines()[-1])
    dd=d["decode"]
    print(*sys.argv[1:], "decode ms", round(dd["ms_per_step"],4), "frac", round(dd["roofline"]["frac"],4), "exact", dd["round_trip_exact"], "| encode frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print(*sys.argv[1:], "failed", e, open('gpurun_out/dec_iter.err').read()[-800:])
PY
done; done; done
