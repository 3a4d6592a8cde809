timeout 400 python bench.py --workload bacterial_rods --cells-per-gpu 2000000 --steps 20 --warmup 5 > gpurun_out/c12_rods2M.json 2> gpurun_out/c12_rods2M.err
tail -1 gpurun_out/c12_rods2M.json | cut -c1-600; tail -3 gpurun_out/c12_rods2M.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload brownian_morse --cells-per-gpu 1000000 --steps 20 --warmup 5 > gpurun_out/c12_brownian_n2.json 2> gpurun_out/c12_brownian_n2.err
tail -1 gpurun_out/c12_brownian_n2.json | cut -c1-300; tail -3 gpurun_out/c12_brownian_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c12_cs_n2.json 2> gpurun_out/c12_cs_n2.err
tail -1 gpurun_out/c12_cs_n2.json | cut -c1-300
