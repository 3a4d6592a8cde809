timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/c18_cs_n8.json 2> gpurun_out/c18_cs_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --workload brownian_morse --cells-per-gpu 1000000 --steps 20 --warmup 5 > gpurun_out/c18_bm_n8.json 2> gpurun_out/c18_bm_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/c18_cs_n4.json 2> gpurun_out/c18_cs_n4.err
for f in gpurun_out/c18_cs_n8.json gpurun_out/c18_bm_n8.json gpurun_out/c18_cs_n4.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d.get('agents_after'))"; done
