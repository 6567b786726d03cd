#!/bin/bash
# Decode-focused GPU visit: decode parity tests, then the decode leg of the bench for each kernel.
# Every step runs under its own short timeout: a deadlocked kernel must not eat the GPU budget.
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_decode_gpu.py -m gpu -x -q --timeout 60 2>&1 | tail -8
for K in ${KERNELS:-pair simd}; do for F in ${FRAMES:-1024 2048}; do
  FPV_DECODE_KERNEL=$K timeout -s KILL 120 python bench.py --steps 20 --warmup 3 --frames $F --no-cpu --no-e2e --no-stream > gpurun_out/bench_dec_${K}_$F.json 2> gpurun_out/bench_dec_${K}_$F.err
  python - "$K" "$F" <<'PY'
import json,sys
K,F=sys.argv[1:3]
try:
    d=json.loads(open(f'gpurun_out/bench_dec_{K}_{F}.json').read().strip().splitlines()[-1])
    dd=d["decode"]; print(K,F,"decode GB/s",round(dd["value"],1),"ms",round(dd["ms_per_step"],4),"exact",dd["round_trip_exact"],"frac",round(dd["roofline"]["frac"],4), "| encode frac", round(d["roofline"]["frac"],4))
except Exception as e:
    print(K,F,"bench failed",e); print(open(f'gpurun_out/bench_dec_{K}_{F}.err').read()[-2000:])
PY
done; done
