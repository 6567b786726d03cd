#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
LIBS="lib lib_all" bash scripts/gpu_abl_enc.sh
WORKLOAD=c3 LIBS="lib lib_all" bash scripts/gpu_abl_enc.sh
WORKLOAD=c1 LIBS="lib" bash scripts/gpu_abl_enc.sh
