mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c7_pytest.log 2>&1
tail -8 gpurun_out/c7_pytest.log
timeout 300 python bench.py --workload bacterial_rods --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/c7_bench_rods.json 2> gpurun_out/c7_bench_rods.err
CRB_ROD_SMEM_PAD=61440 timeout 300 python bench.py --workload bacterial_rods --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/c7_bench_rods_pad60k.json 2> gpurun_out/c7_bench_rods_pad60k.err
timeout 300 python bench.py --workload brownian_morse --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/c7_bench_bm8M.json 2> gpurun_out/c7_bench_bm8M.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/c7_bench8M.json 2> gpurun_out/c7_bench8M.err
for f in gpurun_out/c7_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches","stage_ms_per_step")}, (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
