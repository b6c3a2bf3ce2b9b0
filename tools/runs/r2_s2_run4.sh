# round 2, session 2, call 4: tiled weight packs (microbench both ways, parity suite, step), fused linear wgrad microbench
mkdir -p gpurun_out
python tools/pack_microbench.py > gpurun_out/s2_pack_tiled.log 2>&1; cat gpurun_out/s2_pack_tiled.log
GLIS_PACK_TILED=0 python tools/pack_microbench.py > gpurun_out/s2_pack_elem.log 2>&1; cat gpurun_out/s2_pack_elem.log
python tools/edge_microbench.py > gpurun_out/s2_edge.log 2>&1; tail -8 gpurun_out/s2_edge.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pack_pytest.log 2>&1; tail -3 gpurun_out/s2_pack_pytest.log
for v in 1 0; do
  GLIS_PACK_TILED=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('GLIS_PACK_TILED=$v  %.4f ms  e2e %.4f ms  %d launches' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['details']['launches_per_iteration']))" | tee -a gpurun_out/s2_pack_bench.log
done
