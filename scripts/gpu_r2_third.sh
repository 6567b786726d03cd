#!/bin/bash
# fused encode kernel: parity, bench; backtrace of the columnar encoder test crash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
( timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-ingest --no-stream > gpurun_out/bench_r2_fused.json 2> gpurun_out/bench_r2_fused.err ); tail -c 800 gpurun_out/bench_r2_fused.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r2_fused.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3), "kernel ms", round(d["roofline"]["kernel_ms"], 4),
          "share", round(d["roofline"]["kernel_share_of_step"], 3), "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"], "flags", d["config"]["flags_histogram"])
    for k, v in (d.get("configs") or {}).items():
        print(k, "enc", round(v["encode"]["roofline"]["frac"], 3), "step", round(v["encode"]["roofline"]["step_frac"], 3), "dec", round(v["decode"]["roofline"]["frac"], 3), v["decode"]["round_trip_exact"])
except Exception as e:
    print("bench parse failed", e)
PY
cuda-gdb -batch -ex run -ex bt -ex "info threads" --args oracle/_ref/bin/gpu_mirror_columnar_batch_encoder_test 2>&1 | tail -40
