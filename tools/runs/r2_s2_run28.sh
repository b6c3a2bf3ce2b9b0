# round 2, session 2, call 28 (2 GPUs): final data-parallel sanity — equivalence test, the driver's 2-GPU bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dp_gpu.py -m gpu -x -q > gpurun_out/s2_final_dp_pytest.log 2>&1; tail -2 gpurun_out/s2_final_dp_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/s2_final_bench2.json 2> gpurun_out/s2_final_bench2.err; tail -1 gpurun_out/s2_final_bench2.json | cut -c1-400; grep -i "warn\|error" gpurun_out/s2_final_bench2.err | head -5
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
