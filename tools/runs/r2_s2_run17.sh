# round 2, session 2, call 17: TPReLU backward with two row groups in flight; suite + step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2_r17_pytest.log 2>&1; tail -3 gpurun_out/s2_r17_pytest.log
for v in 1 2; do
  timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4f ms  e2e %.4f ms  %d launches' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['details']['launches_per_iteration']))" | tee -a gpurun_out/s2_r17_bench.log
done
