#!/bin/bash
# ncu evidence for profiles/r02_*: launch list of one bench command + full captures of the dominant kernels.
mkdir -p gpurun_out
K='regex:k_(encode|decode|decide|finalize|gen|delta|cg|combine|planes|entropy)'
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-stream --no-configs --no-ingest"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 160 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_encode_fast -s 9 -c 1 -o gpurun_out/prof_encode_main -f $CMD --no-decode --no-entropy > gpurun_out/ncu_encode.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_decode_fused -s 2 -c 1 -o gpurun_out/prof_decode -f $CMD --no-entropy > gpurun_out/ncu_decode.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_entropy_chunk -s 1 -c 1 -o gpurun_out/prof_entropy -f $CMD --no-decode > gpurun_out/ncu_entropy.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_decode_fused -s 2 -c 1 -o gpurun_out/prof_decode_c3 -f $CMD --workload c3 --no-entropy > gpurun_out/ncu_decode_c3.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_encode_fast -s 9 -c 1 -o gpurun_out/prof_encode_c3 -f $CMD --workload c3 --no-decode --no-entropy > gpurun_out/ncu_encode_c3.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_decode_fused -s 2 -c 1 -o gpurun_out/prof_decode_c1 -f $CMD --workload c1 --no-entropy > gpurun_out/ncu_decode_c1.log 2>&1
python scripts/gpu_entdec.py 64 2>&1 | tail -1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_entropy_decode -s 2 -c 1 -o gpurun_out/prof_entropy_decode -f python scripts/gpu_entdec.py 64 > gpurun_out/ncu_entdec.log 2>&1
tail -2 gpurun_out/ncu_entdec.log
ls -la gpurun_out/*.ncu-rep
# the reports are too large to travel back (gpurun_out is capped at 64 MiB): summarise them here, keep the summaries
python profiles/make_profiles.py r02 2>&1 | tail -3
mkdir -p gpurun_out/profiles_r02
cp profiles/r02_* profiles/traffic.json gpurun_out/profiles_r02/
rm -f gpurun_out/prof_decode.ncu-rep gpurun_out/prof_decode_c1.ncu-rep gpurun_out/prof_decode_c3.ncu-rep gpurun_out/prof_encode_c3.ncu-rep gpurun_out/prof_encode_main.ncu-rep
ls gpurun_out/profiles_r02
