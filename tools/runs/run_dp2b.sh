mkdir -p gpurun_out
GLIS_OVERLAP_WGRAD=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dp_equivalence_gpu.py > gpurun_out/dp2_equiv_nooverlap.log 2>&1; grep -E "dp equivalence|AssertionError" gpurun_out/dp2_equiv_nooverlap.log | head -5
