mkdir -p gpurun_out
nvidia-smi -L | head -3
( time timeout 600 python -m pytest tests/test_gpu_nccl_slabs.py tests/test_gpu_slabs.py -m gpu -x -q ) > gpurun_out/c15_pytest_slabs.log 2>&1
tail -5 gpurun_out/c15_pytest_slabs.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c15_bench_n2_8M.json 2> gpurun_out/c15_bench_n2_8M.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 --workload brownian_morse --cells-per-gpu 4000000 --no-e2e > gpurun_out/c15_bench_n2_bm8M.json 2> gpurun_out/c15_bench_n2_bm8M.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/c15_bench_n2_reference.json 2> gpurun_out/c15_bench_n2_reference.err
for f in gpurun_out/c15_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches","n_gpus","stage_ms_per_step")}, (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-2500:])
PY
done
