#!/bin/bash
# compute-sanitizer memcheck over the kernels that are new in round 2: k_decode_fused (both store paths), k_decode_pair
# in S form, k_decode_spec planes mode, k_entropy_chunk with directories, k_entropy_decode, the fused encode decisions.
mkdir -p gpurun_out
timeout -s KILL 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_r2a.log \
  python -m pytest tests/test_entropy_gpu.py -m gpu -x -q -k "decoder or degenerate or 256-64 or 320" > gpurun_out/memcheck_r2a_pytest.log 2>&1
echo "memcheck entropy rc=$?"; tail -2 gpurun_out/memcheck_r2a_pytest.log; tail -3 gpurun_out/memcheck_r2a.log
timeout -s KILL 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_r2b.log \
  python -m pytest tests/test_decode_gpu.py tests/test_encode_gpu.py -m gpu -x -q -k "golden or every_kernel_mixed_flags and (fused or pair) and (64-5 or 1040 or 1280-2 or 1312 or 2048-1 or 320) or unpredict or device_pointer or without_delta" > gpurun_out/memcheck_r2b_pytest.log 2>&1
echo "memcheck decode/encode rc=$?"; tail -2 gpurun_out/memcheck_r2b_pytest.log; tail -3 gpurun_out/memcheck_r2b.log
