mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 100 --warmup 20 > gpurun_out/bench_8gpu.log 2>gpurun_out/bench_8gpu.err; tail -1 gpurun_out/bench_8gpu.log | cut -c1-400; tail -3 gpurun_out/bench_8gpu.err | cut -c1-300
