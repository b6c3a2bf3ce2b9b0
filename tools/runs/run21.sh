mkdir -p gpurun_out
echo "== pixel-major v8" > gpurun_out/pmmb2.log; timeout 120 python tools/pm_microbench.py >> gpurun_out/pmmb2.log 2>&1
cat gpurun_out/pmmb2.log | cut -c1-200
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/gputest29.log 2>&1; tail -8 gpurun_out/gputest29.log | cut -c1-300
python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench24.log 2>gpurun_out/bench24.err; cut -c1-330 gpurun_out/bench24.log; tail -3 gpurun_out/bench24.err
