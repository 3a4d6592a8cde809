# round-end verification on one B200 (gpurun -- 'bash tools/_call.sh'): gpu tests, smoke, the headline bench line
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1
tail -5 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
