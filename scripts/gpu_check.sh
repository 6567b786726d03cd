#!/bin/bash
# One GPU-box visit: smoke, parity tests, both bench arms.  Run under gpurun.
mkdir -p gpurun_out
nvidia-smi -L; nproc; lscpu | grep "Model name"
python __graft_entry__.py smoke 2>&1 | tail -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
( time python bench.py --steps 50 --warmup 5 > gpurun_out/bench_c2.json ) 2>&1 | tail -3; true 2> gpurun_out/bench_c2.err; tail -c 3000 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
ls -la gpurun_out
