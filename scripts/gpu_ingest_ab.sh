for Z in 0 1 0 1; do
if [ $Z = 1 ]; then export FPV_NO_ZERO_COPY=1; else unset FPV_NO_ZERO_COPY; fi
python bench.py --steps 3 --warmup 3 --no-cpu --no-stream --no-decode --no-entropy --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); g=d['ingest']['gpu_entropy']
print('no_zero_copy=$Z', 'max zero-drop', g['max_zero_drop_fps'], [(int(s['offered_fps']), s['dropped'], round(s['max_ms'],1)) for s in g['sweep']])"
done
