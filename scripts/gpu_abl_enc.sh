#!/bin/bash
# encode-kernel ablation builds (timing only; outputs of the ABL builds are wrong by construction) + one ncu capture
mkdir -p gpurun_out
for L in ${LIBS:-lib lib_np lib_h1 lib_nph1 lib_all}; do
  FPV_B200_LIB=$PWD/fusion_power_video_b200/$L/libfpv_b200.so timeout -s KILL 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-decode --no-cpu --no-stream --no-entropy --no-configs --no-ingest ${WORKLOAD:+--workload $WORKLOAD} > gpurun_out/abl.json 2> gpurun_out/abl.err
  python - $L <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/abl.json').read().strip().splitlines()[-1])
    r=d["roofline"]
    print(sys.argv[1], "step ms", round(d["ms_per_step"],4), "kernel ms", round(r["kernel_ms"],4), "frac", round(r["frac"],3), "launches", d["gpu_launches"])
except Exception as e:
    print(sys.argv[1], "failed", e, open('gpurun_out/abl.err').read()[-500:])
PY
done
if [ -n "$NCU" ]; then
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_encode_fast -s 3 -c 1 -o gpurun_out/prof_encode_fused -f python bench.py --steps 3 --warmup 3 --no-e2e --no-decode --no-cpu --no-stream --no-entropy --no-configs --no-ingest > gpurun_out/ncu_encode_fused.log 2>&1
tail -2 gpurun_out/ncu_encode_fused.log
fi
