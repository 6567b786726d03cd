#!/bin/bash
# whole-codec (stream) leg with eager and with lazy batch allocation in Encoder::Init
for Z in 0 1; do
if [ $Z = 1 ]; then export FPV_LAZY_BATCHES=1; else unset FPV_LAZY_BATCHES; fi
python bench.py --steps 3 --warmup 3 --no-cpu --no-decode --no-entropy --no-configs --no-ingest 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stream']
print('lazy=$Z', 'stream', round(s['value'],2), 'gpu_entropy', round(s['gpu_entropy']['value'],2), 'pinned', round(s['gpu_entropy']['pinned_input']['value'],2))"
done
