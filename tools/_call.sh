timeout 600 python -m pytest tests/test_gpu_rods.py -x -q 2>&1 | tail -3
timeout 400 python bench.py --workload bacterial_rods --cells-per-gpu 2000000 --steps 20 --warmup 5 > gpurun_out/c21_rods2M.json 2> gpurun_out/c21_rods2M.err
python -c "
import json
d=json.loads(open('gpurun_out/c21_rods2M.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['rods_after'])"
