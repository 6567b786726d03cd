#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on the entropy coder) over a slice of the GPU tests.
mkdir -p gpurun_out
timeout -s KILL 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log \
  python -m pytest tests/test_entropy_gpu.py tests/test_cli_gpu.py -m gpu -x -q -k "not full_size" > gpurun_out/memcheck_pytest.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck_pytest.log; grep -c "Invalid\|Error" gpurun_out/memcheck.log; tail -5 gpurun_out/memcheck.log
timeout -s KILL 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck2.log \
  python -m pytest tests/test_encode_gpu.py tests/test_decode_gpu.py -m gpu -x -q -k "golden or narrow or w12" > gpurun_out/memcheck2_pytest.log 2>&1
echo "memcheck2 rc=$?"; tail -3 gpurun_out/memcheck2_pytest.log; tail -4 gpurun_out/memcheck2.log
timeout -s KILL 600 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck.log \
  python -m pytest tests/test_entropy_gpu.py -m gpu -x -q -k "degenerate or 256-64" > gpurun_out/racecheck_pytest.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/racecheck_pytest.log; tail -6 gpurun_out/racecheck.log
