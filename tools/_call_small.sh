mkdir -p gpurun_out
for k in tile2 tile agent; do
CRB_FORCE_KERNEL=$k timeout 120 python bench.py --workload cell_sorting_as_shipped --steps 2000 --warmup 50 --no-cpu-baseline --no-e2e > gpurun_out/c24_as_shipped_$k.json 2>/dev/null
for n in 20000 100000; do
CRB_FORCE_KERNEL=$k timeout 120 python bench.py --cells-per-gpu $n --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/c24_n${n}_$k.json 2>/dev/null
done; done
for f in gpurun_out/c24_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["roofline"]["kernel"][:24], "%.4g agent-steps/s" % d["value"], "%.1f us/step" % (d["ms_per_step"]*1e3), {k:round(v*1e3,1) for k,v in d["stage_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
