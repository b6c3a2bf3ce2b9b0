python -m pytest tests -m gpu -x -q > gpurun_out/gputest11.log 2>&1; tail -5 gpurun_out/gputest11.log
python tools/edge_microbench.py > gpurun_out/edge11.log 2>&1; cat gpurun_out/edge11.log
python tools/wgrad_microbench.py > gpurun_out/wgmb11.log 2>&1; cat gpurun_out/wgmb11.log
python bench.py --steps 50 --warmup 10 > gpurun_out/bench8.log 2>gpurun_out/bench8.err; cut -c1-400 gpurun_out/bench8.log
