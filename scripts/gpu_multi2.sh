#!/bin/bash
# N-GPU visit (N = $1): two-device tests, then bench.py under torchrun like the driver runs it
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8; nproc
timeout -s KILL 300 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -3
( time timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/bench_r2_n$N.err | tail -1 > gpurun_out/bench_r2_n$N.json ) 2>&1 | grep real
tail -3 gpurun_out/bench_r2_n$N.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads(open(f"gpurun_out/bench_r2_n{N}.json").read().strip().splitlines()[-1])
print(f"x{N}: value", round(d["value"],1), "GB/s", round(d["frames_per_s"]), "fps; frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1),
      "of ceiling", round(d["e2e"]["frac_of_pcie_ceiling"],3), {k: round(v,1) for k,v in d["e2e"]["pcie_ceiling"].items() if k.endswith("gbs")},
      "decode", round(d["decode"]["value"],1), "frac", round(d["decode"]["roofline"]["frac"],3), "dec e2e", round(d["decode_e2e"]["value"],1), d["clocks"])
print("stream", d["stream"]["value"], "gpu_entropy", d["stream"]["gpu_entropy"]["value"], "merged ok", d["multi_gpu"]["merged_stream"]["matches_single_gpu_stream"])
PY
