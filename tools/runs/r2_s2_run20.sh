# round 2, session 2, call 20 (2 GPUs): NCCL channel count / protocol for the exposed tail all-reduce
mkdir -p gpurun_out
bash tools/dp_bench.sh 2 NCCL_MIN_NCHANNELS=16 2>&1 | tee -a gpurun_out/s2_nccl_dp2.log
bash tools/dp_bench.sh 2 NCCL_MIN_NCHANNELS=32 2>&1 | tee -a gpurun_out/s2_nccl_dp2.log
bash tools/dp_bench.sh 2 NCCL_MAX_NCHANNELS=4 2>&1 | tee -a gpurun_out/s2_nccl_dp2.log
bash tools/dp_bench.sh 2 NCCL_PROTO=LL128 2>&1 | tee -a gpurun_out/s2_nccl_dp2.log
bash tools/dp_bench.sh 2 NCCL_ALGO=Tree 2>&1 | tee -a gpurun_out/s2_nccl_dp2.log
bash tools/dp_bench.sh 2 GLIS_DP_BUCKET_MB=16 2>&1 | tee -a gpurun_out/s2_nccl_dp2.log
bash tools/dp_bench.sh 2 X=1 2>&1 | tee -a gpurun_out/s2_nccl_dp2.log
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "NCCL INFO (Channel|Connected|comm|NVLS|Using|Ring|Trees|threadThresholds|[0-9]+ coll channels)" | head -30 > gpurun_out/s2_nccl_info.log; head -20 gpurun_out/s2_nccl_info.log
