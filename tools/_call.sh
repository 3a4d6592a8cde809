mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c3_pytest.log 2>&1
tail -8 gpurun_out/c3_pytest.log
timeout 300 python bench.py --workload cell_sorting_as_shipped --steps 2000 --warmup 50 --no-cpu-baseline > gpurun_out/c3_bench_as_shipped.json 2> gpurun_out/c3_bench_as_shipped.err
CRB_NO_GRAPH=1 timeout 300 python bench.py --workload cell_sorting_as_shipped --steps 2000 --warmup 50 --no-cpu-baseline --no-e2e > gpurun_out/c3_bench_as_shipped_nograph.json 2> gpurun_out/c3_bench_as_shipped_nograph.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/c3_bench8M.json 2> gpurun_out/c3_bench8M.err
timeout 300 python bench.py --workload bacterial_rods --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/c3_bench_rods_tight.json 2> gpurun_out/c3_bench_rods_tight.err
CRB_ROD_TIGHT=0 timeout 300 python bench.py --workload bacterial_rods --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/c3_bench_rods_loose.json 2> gpurun_out/c3_bench_rods_loose.err
for f in gpurun_out/c3_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches","stage_ms_per_step")}, (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
