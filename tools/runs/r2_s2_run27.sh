# round 2, session 2, call 27: SMs left free by the persistent kernels' plans, single GPU (the side stream's kernels get them)
mkdir -p gpurun_out
for r in 0 16 8 12 20 0 16; do
  GLIS_RESERVE_SMS=$r timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('GLIS_RESERVE_SMS=$r  %.4f ms  e2e %.4f ms' % (d['ms_per_step'], d['e2e']['ms_per_step']))" | tee -a gpurun_out/s2_reserve_1gpu.log
done
