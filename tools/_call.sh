timeout 600 python -m pytest tests/test_gpu_slabs.py tests/test_gpu_nccl_slabs.py -x -q 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/c17_cs_n2.json 2> gpurun_out/c17_cs_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/c17_cs_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['frac'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --cells-per-gpu 1000000 > gpurun_out/c17_cs_n2_1M.json 2> gpurun_out/c17_cs_n2_1M.err
python -c "
import json
d=json.loads(open('gpurun_out/c17_cs_n2_1M.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['frac'])"
