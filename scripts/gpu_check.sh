#!/bin/bash
# One GPU-box visit: parity tests, bench, launch list, ncu captures.  Run under gpurun.
mkdir -p gpurun_out
nvidia-smi -L; nproc; lscpu | grep "Model name"
python __graft_entry__.py smoke 2>&1 | tail -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 3000 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
K='regex:k_(encode|decode|decide|finalize|gen|delta|cg|combine|planes)'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 3 --warmup 3 --frames 512 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_encode_fast -s 4 -c 2 -o gpurun_out/prof_encode -f \
  python bench.py --steps 3 --warmup 3 --frames 512 --no-e2e --no-cpu --no-decode > gpurun_out/ncu_encode.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_decode_spec -s 1 -c 1 -o gpurun_out/prof_decode -f \
  python bench.py --steps 3 --warmup 3 --frames 512 --no-e2e --no-cpu > gpurun_out/ncu_decode.log 2>&1
ls -la gpurun_out
