mkdir -p gpurun_out
timeout 300 python tools/flaky_config4.py 40 > gpurun_out/flaky2.log 2>&1; tail -6 gpurun_out/flaky2.log | cut -c1-220
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/gputest36.log 2>&1; tail -5 gpurun_out/gputest36.log | cut -c1-300
python __graft_entry__.py smoke > gpurun_out/smoke2.log 2>&1; tail -2 gpurun_out/smoke2.log | cut -c1-250
python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench30.log 2>gpurun_out/bench30.err; cut -c100-260 gpurun_out/bench30.log; tail -2 gpurun_out/bench30.err
