# round 2, session 2, call 18: programmatic dependent launch for SMALL kernels only (mode 1) vs off vs all
mkdir -p gpurun_out
for v in 1 0 2 1 0; do
  GLIS_PDL=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('GLIS_PDL=$v  %.4f ms  e2e %.4f ms  %d launches' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['details']['launches_per_iteration']))" | tee -a gpurun_out/s2_r18_bench.log
done
GLIS_PDL=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2_r18_pytest.log 2>&1; tail -3 gpurun_out/s2_r18_pytest.log
