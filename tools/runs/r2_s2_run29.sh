# round 2, session 2, call 29: final single-GPU validation after the host-side clean-up (mutex-guarded kernel attributes)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2_final2_pytest.log 2>&1; tail -3 gpurun_out/s2_final2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s2_final2_smoke.log 2>&1; tail -1 gpurun_out/s2_final2_smoke.log | cut -c1-300
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4f ms  e2e %.4f ms  %d launches' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['details']['launches_per_iteration']))"
