timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/c8_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force_tile2 -s 3 -c 1 -o gpurun_out/c8_tile2 -f python bench.py --steps 2 --warmup 3 --cells-per-gpu 1000000 --no-cpu-baseline --no-e2e > gpurun_out/c8_ncu.log 2>&1
python tools/ncu_regions.py gpurun_out/c8_tile2.ncu-rep tile2 > gpurun_out/c8_regions.txt 2>&1
python tools/ncu_summary.py gpurun_out/c8_tile2.ncu-rep 40 > gpurun_out/c8_summary.txt 2>&1
cat gpurun_out/c8_pytest.log gpurun_out/c8_regions.txt
