# round 2, session 2, call 19 (2 GPUs): SMs reserved for NCCL under data parallelism
mkdir -p gpurun_out
for r in 0 8 16 24 0; do
  bash tools/dp_bench.sh 2 GLIS_RESERVE_SMS=$r 2>&1 | tee -a gpurun_out/s2_reserve_dp2.log
done
bash tools/dp_bench.sh 1 GLIS_RESERVE_SMS=0 2>&1 | tee -a gpurun_out/s2_reserve_dp2.log
bash tools/dp_bench.sh 1 GLIS_RESERVE_SMS=16 2>&1 | tee -a gpurun_out/s2_reserve_dp2.log
