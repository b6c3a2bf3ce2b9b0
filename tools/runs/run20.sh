mkdir -p gpurun_out
echo "== pixel-major" > gpurun_out/pmmb.log; timeout 120 python tools/pm_microbench.py >> gpurun_out/pmmb.log 2>&1
echo "== channel-major (GLIS_TC_PM=0)" >> gpurun_out/pmmb.log; GLIS_TC_PM=0 timeout 120 python tools/pm_microbench.py >> gpurun_out/pmmb.log 2>&1
cat gpurun_out/pmmb.log | cut -c1-200
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/gputest28.log 2>&1; tail -8 gpurun_out/gputest28.log | cut -c1-300
python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench23.log 2>gpurun_out/bench23.err; cut -c1-330 gpurun_out/bench23.log; tail -3 gpurun_out/bench23.err
