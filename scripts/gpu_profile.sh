#!/bin/bash
# ncu evidence for profiles/: launch list of one bench command + full captures of the dominant kernels.
mkdir -p gpurun_out
K='regex:k_(encode|decode|decide|finalize|gen|delta|cg|combine|planes|entropy)'
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-stream"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_encode_fast -s 3 -c 1 -o gpurun_out/prof_encode_main -f $CMD --no-decode --no-entropy > gpurun_out/ncu_encode.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_decode_pair -s 1 -c 1 -o gpurun_out/prof_decode -f $CMD --no-entropy > gpurun_out/ncu_decode.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_entropy_chunk -s 1 -c 1 -o gpurun_out/prof_entropy -f $CMD --no-decode > gpurun_out/ncu_entropy.log 2>&1
ls -la gpurun_out | head -30
