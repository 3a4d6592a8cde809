# one full ncu capture of the dominant kernel of the headline and of the Brownian+Morse workload at 8 M cells;
# only the text summaries travel back (the .ncu-rep files are ~46 MB each)
mkdir -p gpurun_out /tmp/rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force_tile2 --launch-skip 4 -c 1 -f -o /tmp/rep/tile2_cs_8M python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c21_a.log 2>&1
python tools/ncu_summary.py /tmp/rep/tile2_cs_8M.ncu-rep 30 > gpurun_out/c21_ncu_k_force_tile2_cs_8M.txt 2>&1
python tools/ncu_regions.py /tmp/rep/tile2_cs_8M.ncu-rep tile2 > gpurun_out/c21_ncu_k_force_tile2_cs_8M_regions.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force_tile2 --launch-skip 4 -c 1 -f -o /tmp/rep/tile2_morse_8M python bench.py --workload brownian_morse --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c21_b.log 2>&1
python tools/ncu_summary.py /tmp/rep/tile2_morse_8M.ncu-rep 30 > gpurun_out/c21_ncu_k_force_tile2_morse_8M.txt 2>&1
python tools/ncu_regions.py /tmp/rep/tile2_morse_8M.ncu-rep tile2 > gpurun_out/c21_ncu_k_force_tile2_morse_8M_regions.txt 2>&1
ls -la gpurun_out
