# round-end verification on one B200 (gpurun -- 'bash tools/_call.sh'): gpu tests, smoke, bench lines
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1
tail -5 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
timeout 200 python bench.py --workload cell_sorting_as_shipped --steps 2000 --warmup 50 > gpurun_out/final_bench_as_shipped.json 2> gpurun_out/final_bench_as_shipped.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/final_bench8M_noe2e.json 2> gpurun_out/final_bench8M_noe2e.err
for f in gpurun_out/final_bench*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["roofline"]["kernel"][:30], "%.4g" % d["value"], "%.2f us/step" % (d["ms_per_step"]*1e3), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
