mkdir -p gpurun_out
cd _wt && timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dp_equivalence_gpu.py > ../gpurun_out/dp2_equiv_old.log 2>&1; grep -E "dp equivalence|AssertionError" ../gpurun_out/dp2_equiv_old.log | head -5
