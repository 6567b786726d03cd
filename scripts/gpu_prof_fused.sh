#!/bin/bash
# full-occupancy timing of the fused decode kernel + one ncu --set full capture of it
mkdir -p gpurun_out
KERNELS=fused FRAMES="${FRAMES:-2368}" WLS="${WLS:-c2 c1 c3}" bash scripts/gpu_dec_frames2.sh
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_decode_fused -s 2 -c 1 -o gpurun_out/prof_decode_fused -f \
  python bench.py --steps 3 --warmup 3 --frames 2368 --decode-frames 2368 --no-e2e --no-cpu --no-stream --no-entropy --no-configs --no-ingest > gpurun_out/ncu_decode_fused.log 2>&1
tail -3 gpurun_out/ncu_decode_fused.log
