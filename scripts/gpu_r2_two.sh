#!/bin/bash
# Round-2 two-GPU visit: the 2-GPU tests, the sharded configs[2] bench under torchrun, the reference arm, PCIe ceilings.
mkdir -p gpurun_out
nvidia-smi -L | head -3; nproc
for t in columnar_batch_decoder_test columnar_batch_encoder_test; do
  timeout 120 oracle/_ref/bin/gpu_mirror_$t > gpurun_out/col_$t.out 2> gpurun_out/col_$t.err; echo "gpu_mirror_$t rc=$?"; tail -3 gpurun_out/col_$t.err; grep -c "Got the Batch" gpurun_out/col_$t.out; grep Closed gpurun_out/col_$t.out
done
timeout -s KILL 400 python -m pytest tests/test_multigpu_gpu.py tests/test_reference_programs.py -m gpu -q 2>&1 | tail -8
( time timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err ) 2>&1 | tail -3
tail -c 2500 gpurun_out/bench_r2_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r2_n2.json").read().strip().splitlines()[-1])
    print("workload", d["config"]["workload"][:60], "scaling", d["scaling"])
    print("value", round(d["value"], 1), "fps", round(d["frames_per_s"]), "ms/step", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3),
          "e2e", round(d["e2e"]["value"], 1), "decode frac", round(d["decode"]["roofline"]["frac"], 3), "launches", d["gpu_launches"])
    print(json.dumps(d["multi_gpu"])[:1500])
    print("stream", d["stream"]["value"], d["stream"]["gpu_entropy"]["value"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
timeout -s KILL 200 python scripts/gpu_pcie_multi.py --seconds 1.0 --out gpurun_out/pcie_multi_2gpu.json 2>&1 | grep gpus
