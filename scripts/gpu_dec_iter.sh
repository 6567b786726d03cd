#!/bin/bash
# decode-kernel iteration: decode parity tests, then the decode leg per workload (c2, c1, c3) for each library in LIBS
mkdir -p gpurun_out
if [ -z "$NOTEST" ]; then timeout -s KILL 300 python -m pytest tests/test_decode_gpu.py -m gpu -x -q 2>&1 | tail -3; fi
for L in ${LIBS:-lib}; do for WL in ${WLS:-c2 c1 c3}; do
  FPV_B200_LIB=$PWD/fusion_power_video_b200/$L/libfpv_b200.so timeout -s KILL 120 python bench.py --steps 20 --warmup 3 --workload $WL --no-e2e --no-cpu --no-stream --no-entropy --no-configs --no-ingest > gpurun_out/dec_iter.json 2> gpurun_out/dec_iter.err
  python - $L $WL <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/dec_iter.json').read().strip().splitlines()[-1])
    dd=d["decode"]
    print(sys.argv[1], sys.argv[2], "decode ms", round(dd["ms_per_step"],4), "frac", round(dd["roofline"]["frac"],4), "exact", dd["round_trip_exact"], "| encode frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "failed", e, open('gpurun_out/dec_iter.err').read()[-800:])
PY
done; done
if [ -n "$PROF" ]; then
  FPV_B200_LIB=$PWD/fusion_power_video_b200/lib_prof/libfpv_b200.so timeout -s KILL 200 python scripts/gpu_pair_prof.py c2 c1 c3 2>&1 | tail -3
fi
