# round 2, session 2, call 7: wide-store packs, 16-byte norm / projection kernels
mkdir -p gpurun_out
python tools/pack_microbench.py > gpurun_out/s2_pack_wide.log 2>&1; cat gpurun_out/s2_pack_wide.log
python tools/project_microbench.py > gpurun_out/s2_project2.log 2>&1; grep -v "accumulate" gpurun_out/s2_project2.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2_r7_pytest.log 2>&1; tail -3 gpurun_out/s2_r7_pytest.log
for v in 1 0; do
  GLIS_PACK_WIDE=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('GLIS_PACK_WIDE=$v  %.4f ms  e2e %.4f ms  %d launches' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['details']['launches_per_iteration']))" | tee -a gpurun_out/s2_r7_bench.log
done
