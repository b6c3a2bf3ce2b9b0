# round 2, session 2, call 15 (2 GPUs): data-parallel equivalence test + 2-GPU step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -x -q > gpurun_out/s2_dp2_pytest.log 2>&1; tail -3 gpurun_out/s2_dp2_pytest.log
bash tools/dp_bench.sh 2 X=1 2>&1 | tee gpurun_out/s2_dp2_bench.log
bash tools/dp_bench.sh 1 X=1 2>&1 | tee -a gpurun_out/s2_dp2_bench.log
