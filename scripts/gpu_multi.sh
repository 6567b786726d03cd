#!/bin/bash
# bench.py under torchrun on N GPUs of one box (N = $1), both arms like the driver runs them.
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 2> gpurun_out/b$N.err | tail -1 > gpurun_out/bench_${N}gpu.json
tail -3 gpurun_out/b$N.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads(open(f"gpurun_out/bench_{N}gpu.json").read().strip().splitlines()[-1])
print(f"x{N}: value", round(d["value"],1), "GB/s", round(d["frames_per_s"]), "fps; frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1),
      "decode", round(d["decode"]["value"],1), "dec e2e", round(d["decode_e2e"]["value"],1), d["clocks"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
