mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/gputest25.log 2>&1; tail -6 gpurun_out/gputest25.log | cut -c1-300
for v in "" "GLIS_TC_AROWS=128" "GLIS_TC_NMAX=160" "GLIS_TC_NMAX=128" "GLIS_TC_DEPTH_PENALTY=40"; do
  echo "== $v" >> gpurun_out/tcmb25.log
  env $v python tools/tc_microbench.py 1 >> gpurun_out/tcmb25.log 2>&1
done
cat gpurun_out/tcmb25.log | cut -c1-120
python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench20.log 2>gpurun_out/bench20.err; cut -c1-330 gpurun_out/bench20.log; tail -3 gpurun_out/bench20.err
