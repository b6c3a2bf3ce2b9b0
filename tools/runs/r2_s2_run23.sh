# round 2, session 2, call 23 (2 GPUs): our peer-memory all-reduce in the gradient exchange — equivalence test, step with / without
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dp_gpu.py -m gpu -x -q > gpurun_out/s2_peer_pytest.log 2>&1; tail -15 gpurun_out/s2_peer_pytest.log
timeout 200 bash tools/dp_bench.sh 2 X=1 2>&1 | tee -a gpurun_out/s2_peer_dp2.log
timeout 200 bash tools/dp_bench.sh 2 GLIS_DP_PEER=0 2>&1 | tee -a gpurun_out/s2_peer_dp2.log
timeout 200 bash tools/dp_bench.sh 2 GLIS_DP_BUCKET_MB=64 2>&1 | tee -a gpurun_out/s2_peer_dp2.log
timeout 200 bash tools/dp_bench.sh 2 GLIS_DP_BUCKET_MB=4 2>&1 | tee -a gpurun_out/s2_peer_dp2.log
