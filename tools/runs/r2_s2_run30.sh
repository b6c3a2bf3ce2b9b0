# round 2, session 2, call 30 (4 GPUs): the iteration at 4 ranks with the peer all-reduce (the WORLD = 4 instantiation)
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 tools/symm_allreduce_probe.py > gpurun_out/s2_peer_probe4.log 2>&1; grep -E "13.1 MB at|FAILED|Error" gpurun_out/s2_peer_probe4.log | head -6
timeout 120 bash tools/dp_bench.sh 4 X=1 2>&1 | tee -a gpurun_out/s2_peer_dp4.log
