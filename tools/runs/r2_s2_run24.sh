# round 2, session 2, call 24 (2 GPUs): our peer all-reduce kernel alone vs NCCL vs torch two-shot
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/symm_allreduce_probe.py > gpurun_out/s2_peer_probe.log 2>&1; grep -E "MB at|FAILED|Error|error" gpurun_out/s2_peer_probe.log | head -20
