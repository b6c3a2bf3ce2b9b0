# round 2, session 2, call 13: cta_group::2 pair kernel inside the step (parity suite with it on, step on / off)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2_r13_pytest.log 2>&1; tail -3 gpurun_out/s2_r13_pytest.log
for v in 1 0 1 0; do
  GLIS_TC_PAIR=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('GLIS_TC_PAIR=$v  %.4f ms  e2e %.4f ms  %d launches' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['details']['launches_per_iteration']))" | tee -a gpurun_out/s2_r13_bench.log
done
