mkdir -p gpurun_out
timeout 100 python bench.py --cells-per-gpu 1000000 --steps 50 --warmup 10 > gpurun_out/c29_bench1M.json 2> gpurun_out/c29_bench1M.err
timeout 60 python bench.py --density 2.5 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/c29_bench8M_rho2p5.json 2> gpurun_out/c29_bench8M_rho2p5.err
for f in gpurun_out/c29_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "%.4g" % d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline_hbm"]["frac"], (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
