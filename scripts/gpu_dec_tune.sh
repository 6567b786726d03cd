#!/bin/bash
# decode leg (default kernel, one full wave) per workload for each tuning build in LIBS; optional ncu capture of one workload
mkdir -p gpurun_out
for L in ${LIBS:-lib}; do for WL in ${WLS:-c2 c1 c3}; do
  FPV_B200_LIB=$PWD/fusion_power_video_b200/$L/libfpv_b200.so timeout -s KILL 120 python bench.py --steps 10 --warmup 3 --workload $WL --no-e2e --no-cpu --no-stream --no-entropy --no-configs --no-ingest > gpurun_out/dec_iter.json 2> gpurun_out/dec_iter.err
  python - $L $WL <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/dec_iter.json').read().strip().splitlines()[-1])
    dd=d["decode"]
    print(*sys.argv[1:], "frames", dd["frames_per_gpu"], "decode ms", round(dd["ms_per_step"],4), "frac", round(dd["roofline"]["frac"],4), "exact", dd["round_trip_exact"], "| encode frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print(*sys.argv[1:], "failed", e, open('gpurun_out/dec_iter.err').read()[-800:])
PY
done; done
if [ -n "$NCU_WL" ]; then
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_decode_fused -s 2 -c 1 -o gpurun_out/prof_decode_fused_$NCU_WL -f \
  python bench.py --steps 3 --warmup 3 --workload $NCU_WL --no-e2e --no-cpu --no-stream --no-entropy --no-configs --no-ingest > gpurun_out/ncu_decode_fused.log 2>&1
tail -2 gpurun_out/ncu_decode_fused.log
fi
