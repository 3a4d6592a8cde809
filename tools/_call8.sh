mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 20 --warmup 5 --workload brownian_morse --cells-per-gpu 1000000 --no-e2e > gpurun_out/c22_bench_n8_bm8M_total.json 2> gpurun_out/c22_bench_n8_bm8M_total.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e > gpurun_out/c22_bench_n8_8M.json 2> gpurun_out/c22_bench_n8_8M.err
for f in gpurun_out/c22_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches","n_gpus","stage_ms_per_step")}, (d.get("roofline") or {}).get("frac"))
except Exception as e:
    print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-2500:])
PY
done
