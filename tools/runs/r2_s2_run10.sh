# round 2, session 2, call 10: MN-major shared-box probe; LIS kernel with both packs staged up front; head GEMV / outer product
mkdir -p gpurun_out
timeout 60 tools/probes/umma_mnmajor_shift_probe > gpurun_out/s2_mnprobe.log 2>&1; cat gpurun_out/s2_mnprobe.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2_r10_pytest.log 2>&1; tail -3 gpurun_out/s2_r10_pytest.log
for v in 1 0; do
  GLIS_HEAD_FAST=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('GLIS_HEAD_FAST=$v  %.4f ms  e2e %.4f ms  %d launches' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['details']['launches_per_iteration']))" | tee -a gpurun_out/s2_r10_bench.log
done
python tools/edge_microbench.py > gpurun_out/s2_edge4.log 2>&1; head -8 gpurun_out/s2_edge4.log
