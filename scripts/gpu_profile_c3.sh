#!/bin/bash
# ncu --set full of the pair decode kernel in split mode (2048x2048, one wave = 592 frames) and of the encode kernel at that width.
mkdir -p gpurun_out
CMD="python bench.py --workload c3 --frames 592 --steps 3 --warmup 3 --no-e2e --no-cpu --no-stream --no-entropy"
ncu --set full --clock-control none --import-source on -k regex:k_decode_pair -s 1 -c 1 -o gpurun_out/prof_decode_c3 -f $CMD > gpurun_out/ncu_decode_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_encode_fast -s 3 -c 1 -o gpurun_out/prof_encode_c3 -f $CMD --no-decode > gpurun_out/ncu_encode_c3.log 2>&1
ls -la gpurun_out/*c3*
