set -x
timeout 600 python -m pytest tests/test_gpu_slabs.py tests/test_gpu_nccl_slabs.py -x -q 2>&1 | tail -5 > gpurun_out/c2_pytest.log
for m in "" "--no-overlap"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --cells-per-gpu 2000000 $m > gpurun_out/c2_bench_n2_2M$m.json 2> gpurun_out/c2_bench_n2$m.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 $m > gpurun_out/c2_bench_n2_8M$m.json 2> gpurun_out/c2_bench_n2_8M$m.err
done
cat gpurun_out/c2_pytest.log
for f in gpurun_out/c2_bench_n2_*.json; do echo $f; tail -1 $f | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['e2e']['value'])"; done
