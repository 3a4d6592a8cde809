timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/c22_bench8M.json 2> gpurun_out/c22_bench8M.err
python -c "
import json
d=json.loads(open('gpurun_out/c22_bench8M.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['frac'], d['roofline_hbm']['frac'])"
