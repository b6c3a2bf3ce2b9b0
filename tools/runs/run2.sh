mkdir -p gpurun_out
python tools/debug_flip.py fp32 > gpurun_out/flip12.log 2>&1; tail -12 gpurun_out/flip12.log
python -m pytest tests -m gpu -q > gpurun_out/gputest12.log 2>&1; tail -5 gpurun_out/gputest12.log
python tools/wgrad_microbench.py 0:128 1:128 0:256 1:256 2:256 4:256 > gpurun_out/wgmb12.log 2>&1; cat gpurun_out/wgmb12.log
python bench.py --steps 50 --warmup 10 > gpurun_out/bench9.log 2>gpurun_out/bench9.err; cut -c1-300 gpurun_out/bench9.log
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 600 --csv --log-file gpurun_out/launches_warm12.csv python bench.py --steps 2 --warmup 3 --no-graph > gpurun_out/ncu12.log 2>&1; tail -1 gpurun_out/ncu12.log | cut -c1-100
