#!/bin/bash
# PCIe ceilings + the e2e (host-buffer) leg of the bench over batch sizes.
mkdir -p gpurun_out
python scripts/gpu_pcie.py 64; python scripts/gpu_pcie.py 16; python scripts/gpu_pcie.py 256
for B in ${BATCHES:-16 32 64 128}; do
  python bench.py --steps 10 --warmup 3 --no-cpu --no-decode --no-stream --e2e-batch $B --e2e-frames 512 > gpurun_out/e2e_$B.json 2> gpurun_out/e2e_$B.err
  python - $B <<'PY'
import json,sys
B=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/e2e_{B}.json').read().strip().splitlines()[-1])
    print("batch",B,"e2e GB/s",round(d["e2e"]["value"],2),"fps",round(d["e2e"]["frames_per_s"]))
except Exception as e:
    print(B,"failed",e,open(f'gpurun_out/e2e_{B}.err').read()[-800:])
PY
done
