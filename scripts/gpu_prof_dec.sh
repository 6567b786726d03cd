#!/bin/bash
# ncu --set full capture of the decode kernel (default: k_decode_pair) inside the bench's decode leg.
mkdir -p gpurun_out
KREGEX=${KREGEX:-k_decode_pair}
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 1 -c 1 -o gpurun_out/prof_decode -f \
  python bench.py --steps 3 --warmup 3 --frames ${FRAMES:-1024} --no-e2e --no-cpu --no-stream > gpurun_out/ncu_decode.log 2>&1
tail -3 gpurun_out/ncu_decode.log; ls -la gpurun_out | head
