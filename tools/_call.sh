timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/c10_bench_n8_8M.json 2> gpurun_out/c10_bench_n8.err
tail -1 gpurun_out/c10_bench_n8_8M.json | cut -c1-400; tail -5 gpurun_out/c10_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/c10_bench_n4_8M.json 2> gpurun_out/c10_bench_n4.err
tail -1 gpurun_out/c10_bench_n4_8M.json | cut -c1-400
