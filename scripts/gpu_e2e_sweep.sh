for CFG in "3 64 512" "4 64 512" "4 32 512" "3 128 768" "4 128 1024" "2 128 512"; do set -- $CFG
python bench.py --steps 10 --warmup 3 --no-decode --no-cpu --no-stream --no-entropy --no-configs --no-ingest --e2e-slots $1 --e2e-batch $2 --e2e-frames $3 > gpurun_out/e2e_sweep.json 2>/dev/null
python - $1 $2 $3 <<'PY'
import json,sys
d=json.loads(open('gpurun_out/e2e_sweep.json').read().strip().splitlines()[-1]); e=d["e2e"]
print("slots batch frames", *sys.argv[1:], "e2e", round(e["value"],2), "ceiling", round(e["pcie_ceiling"]["bidir_each_sum_gbs"],1), "frac", round(e["frac_of_pcie_ceiling"],3))
PY
done
