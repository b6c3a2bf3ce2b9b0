# round 2, session 2, call 26: final single-GPU validation — suite, smoke, the driver's bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2_final_pytest.log 2>&1; tail -3 gpurun_out/s2_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s2_final_smoke.log 2>&1; tail -2 gpurun_out/s2_final_smoke.log
timeout 400 python bench.py > gpurun_out/s2_final_bench.json 2> gpurun_out/s2_final_bench.err; cut -c1-900 gpurun_out/s2_final_bench.json
