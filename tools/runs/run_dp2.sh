mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dp_equivalence_gpu.py > gpurun_out/dp2_equiv.log 2>&1; tail -8 gpurun_out/dp2_equiv.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/bench_2gpu_b.log 2>gpurun_out/bench_2gpu_b.err; cut -c1-400 gpurun_out/bench_2gpu_b.log; tail -3 gpurun_out/bench_2gpu_b.err
