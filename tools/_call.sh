set -x
export CRB_DEBUG_PLAN=1
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -15 > gpurun_out/c4_pytest_k.log
timeout 300 python bench.py --steps 20 --warmup 5 --cells-per-gpu 1000000 --no-cpu-baseline --no-e2e > gpurun_out/c4_bench1M.json 2> gpurun_out/c4_bench1M.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/c4_bench8M.json 2> gpurun_out/c4_bench8M.err
CRB_FORCE_KERNEL=tile timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/c4_bench8M_v7.json 2> gpurun_out/c4_bench8M_v7.err
unset CRB_DEBUG_PLAN
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/c4_pytest.log
cat gpurun_out/c4_pytest_k.log;  tail -3 gpurun_out/c4_pytest.log; grep -h "tile2 plan" gpurun_out/c4_bench8M.err | head -2
for f in gpurun_out/c4_bench1M.json gpurun_out/c4_bench8M.json gpurun_out/c4_bench8M_v7.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['frac'])"; done
