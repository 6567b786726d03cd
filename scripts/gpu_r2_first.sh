#!/bin/bash
# Round-2 first visit: parity suite, per-role cycle accounting of the decode kernel, PCIe ceilings + topology.
mkdir -p gpurun_out
nvidia-smi -L; nproc; lscpu | grep -E "Model name|NUMA|Socket"
timeout -s KILL 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
FPV_B200_LIB=$PWD/fusion_power_video_b200/lib_prof/libfpv_b200.so timeout -s KILL 200 python scripts/gpu_pair_prof.py c2 c1 c3 2>&1 | tail -5
timeout -s KILL 200 python scripts/gpu_pcie_multi.py --seconds 1.0 2>&1 | tail -30
