mkdir -p gpurun_out
timeout 300 python bench.py --workload bacterial_rods --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/c10_bench_rods_split_b4.json 2> gpurun_out/c10_bench_rods_split_b4.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/c10_launches_rods_split.csv python bench.py --workload bacterial_rods --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
grep -E "k_rod_force|k_rod_update" gpurun_out/c10_launches_rods_split.csv | awk -F'","' '{print $5, $(NF)}' | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rod_force --launch-skip 3 -c 1 -f -o gpurun_out/c10_rod_force_only_2M python bench.py --workload bacterial_rods --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c10_ncu_full.log 2>&1
for f in gpurun_out/c10_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches","stage_ms_per_step")}, (d.get("roofline") or {}).get("frac"))
except Exception as e:
    print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
