timeout 900 python -m pytest tests/test_gpu_backend.py -x -q 2>&1 | tail -4 > gpurun_out/c9_pytest.log
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_force_tile2|k_reorder|k_fill" -s 12 -c 6 --csv --log-file gpurun_out/c9_traffic8M.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c9_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c9_launches8M.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c9_ncu2.log 2>&1
timeout 600 python bench.py > gpurun_out/c9_bench8M.json 2> gpurun_out/c9_bench8M.err
timeout 600 python bench.py --cells-per-gpu 1000000 > gpurun_out/c9_bench1M.json 2> gpurun_out/c9_bench1M.err
cat gpurun_out/c9_pytest.log; tail -8 gpurun_out/c9_traffic8M.csv; tail -1 gpurun_out/c9_bench8M.json | cut -c1-300
