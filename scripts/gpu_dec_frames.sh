#!/bin/bash
# Decode rate vs frames in flight (C2): the chain bounds a frame at ~1 ms, so the rate grows with the number of frames until a wave is full.
mkdir -p gpurun_out
for F in 4 64 148 296 592 1024 1184 2368 4736; do
  python bench.py --frames $F --steps 10 --warmup 3 --no-cpu --no-e2e --no-stream --no-entropy > gpurun_out/df.json 2> gpurun_out/df.err
  python - $F <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/df.json').read().strip().splitlines()[-1])
    print(sys.argv[1], "frames: decode ms", round(d["decode"]["ms_per_step"],4), "GB/s", round(d["decode"]["value"],1), "frac", round(d["decode"]["roofline"]["frac"],3), "| encode GB/s", round(d["value"],1), "frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print(sys.argv[1], "failed", e, open('gpurun_out/df.err').read()[-300:])
PY
done
