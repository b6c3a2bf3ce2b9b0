# round 2, session 2, call 25 (8 GPUs): peer all-reduce kernel vs NCCL at 8 ranks (probe: agreement + time), iteration with / without
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/symm_allreduce_probe.py > gpurun_out/s2_peer_probe8.log 2>&1; grep -E "MB at|FAILED|Error|error" gpurun_out/s2_peer_probe8.log | head -12
timeout 120 bash tools/dp_bench.sh 8 X=1 2>&1 | tee -a gpurun_out/s2_peer_dp8.log
timeout 120 bash tools/dp_bench.sh 8 GLIS_DP_PEER=0 2>&1 | tee -a gpurun_out/s2_peer_dp8.log
