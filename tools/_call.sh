# round-end verification on one B200 (gpurun -- 'bash tools/_call.sh'): gpu tests, smoke, the three bench lines
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1
tail -5 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench8M.json 2> gpurun_out/final_bench8M.err
timeout 400 python bench.py --workload bacterial_rods --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/final_bench_rods2M.json 2> gpurun_out/final_bench_rods2M.err
for f in gpurun_out/final_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches","stage_ms_per_step")}, (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
