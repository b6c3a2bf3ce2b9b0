mkdir -p gpurun_out
python tools/wgrad_microbench.py 0:256:0 0:256:4 0:256:10 0:256:20 0:256:40 1:256:10 > gpurun_out/wgmb18.log 2>&1; cat gpurun_out/wgmb18.log
python -m pytest tests -m gpu -q -x > gpurun_out/gputest18.log 2>&1; tail -3 gpurun_out/gputest18.log | cut -c1-250
python bench.py --steps 50 --warmup 10 > gpurun_out/bench14.log 2>gpurun_out/bench14.err; cut -c1-300 gpurun_out/bench14.log; tail -3 gpurun_out/bench14.err
