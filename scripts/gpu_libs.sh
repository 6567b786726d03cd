#!/bin/bash
# Decode leg of the bench for alternative builds of libfpv_b200.so (FPV_B200_LIB), per workload.
mkdir -p gpurun_out
for WL in ${WORKLOADS:-c1 c2}; do for L in ${LIBS:-lib lib_k8 lib_k24 lib_k32 lib_g4}; do
  FPV_B200_LIB=$PWD/fusion_power_video_b200/$L/libfpv_b200.so python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu --no-e2e --no-stream --no-entropy > gpurun_out/lib.json 2> gpurun_out/lib.err
  python - $WL $L <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/lib.json').read().strip().splitlines()[-1])
    print(sys.argv[1], sys.argv[2], "decode ms", round(d["decode"]["ms_per_step"],4), "frac", round(d["decode"]["roofline"]["frac"],3), "exact", d["decode"]["round_trip_exact"], "| enc frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "failed", e, open('gpurun_out/lib.err').read()[-600:])
PY
done; done
