mkdir -p gpurun_out
GLIS_DP_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/debug_dp_div.py 0 > gpurun_out/dp2_dbg.log 2>&1; grep -E "^\[dp\]" gpurun_out/dp2_dbg.log | head -70
