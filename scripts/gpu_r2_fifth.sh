#!/bin/bash
# encode: parity, timing per geometry, ncu launch list + full capture (profiles/r02_*)
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_encode_gpu.py tests/test_host_gpu.py -m gpu -x -q 2>&1 | tail -3
LIBS="lib" bash scripts/gpu_abl_enc.sh
WORKLOAD=c3 LIBS="lib" bash scripts/gpu_abl_enc.sh
WORKLOAD=c1 LIBS="lib" bash scripts/gpu_abl_enc.sh
K='regex:k_(encode|decode|decide|finalize|gen|delta|cg|combine|planes|entropy|split)'
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-stream --no-configs --no-ingest"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_launches.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_encode_fast -s 9 -c 1 -o gpurun_out/prof_encode_r02 -f $CMD --no-decode --no-entropy > gpurun_out/ncu_encode.log 2>&1
tail -2 gpurun_out/ncu_encode.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_r02.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows[-14:]:
    print(r[4][:60], r[-1])
PY
