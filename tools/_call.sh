mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_rods.py tests/test_gpu_backend.py -m gpu -x -q ) > gpurun_out/c14_pytest.log 2>&1
tail -5 gpurun_out/c14_pytest.log
timeout 300 python bench.py --workload bacterial_rods --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/c14_bench_rods_bal.json 2> gpurun_out/c14_bench_rods_bal.err
CRB_ROD_BALANCED=0 timeout 300 python bench.py --workload bacterial_rods --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/c14_bench_rods_unbal.json 2> gpurun_out/c14_bench_rods_unbal.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/c14_launches_rods.csv python bench.py --workload bacterial_rods --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
grep -E "k_rod_pairs|k_rod_update" gpurun_out/c14_launches_rods.csv | awk -F'","' '{print $5, $(NF)}' | tail -4
for f in gpurun_out/c14_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches","stage_ms_per_step")}, (d.get("roofline") or {}).get("frac"))
except Exception as e:
    print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
