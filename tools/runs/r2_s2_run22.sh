# round 2, session 2, call 22 (2 GPUs): symmetric-memory all-reduce probe
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/symm_allreduce_probe.py > gpurun_out/s2_symm_probe.log 2>&1; tail -25 gpurun_out/s2_symm_probe.log
