# round 2, session 2, call 21 (8 GPUs): the iteration at 8 GPUs — default NCCL settings vs 32 channels vs 32 channels + 16 MB buckets
mkdir -p gpurun_out
bash tools/dp_bench.sh 8 X=1 2>&1 | tee -a gpurun_out/s2_dp8.log
bash tools/dp_bench.sh 8 NCCL_MIN_NCHANNELS=32 2>&1 | tee -a gpurun_out/s2_dp8.log
bash tools/dp_bench.sh 8 NCCL_MIN_NCHANNELS=32 GLIS_DP_BUCKET_MB=16 2>&1 | tee -a gpurun_out/s2_dp8.log
bash tools/dp_bench.sh 8 GLIS_DP_BUCKET_MB=16 2>&1 | tee -a gpurun_out/s2_dp8.log
