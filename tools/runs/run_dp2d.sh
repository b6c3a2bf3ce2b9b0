mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/debug_dp_div.py 01 > gpurun_out/dp2_div.log 2>&1; grep -E "graph=" gpurun_out/dp2_div.log | head -60
