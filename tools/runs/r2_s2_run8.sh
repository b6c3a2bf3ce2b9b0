# round 2, session 2, call 8: linear-layer pack paths (octets), LIS-sized fused linear wgrad blocks; fresh warm launch list
mkdir -p gpurun_out
python tools/pack_microbench.py > gpurun_out/s2_pack_wide2.log 2>&1; tail -2 gpurun_out/s2_pack_wide2.log
python tools/edge_microbench.py > gpurun_out/s2_edge3.log 2>&1; tail -6 gpurun_out/s2_edge3.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2_r8_pytest.log 2>&1; tail -3 gpurun_out/s2_r8_pytest.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4f ms  e2e %.4f ms  %d launches' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['details']['launches_per_iteration']))" | tee -a gpurun_out/s2_r8_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1200 --csv --log-file gpurun_out/s2_launches_warm2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/s2_ncu3.log 2>&1
