mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/gputest14.log 2>&1; tail -4 gpurun_out/gputest14.log
python tools/edge_microbench.py > gpurun_out/edge14.log 2>&1; grep -i "linear\|head\|LIS" gpurun_out/edge14.log
python bench.py --steps 50 --warmup 10 > gpurun_out/bench11.log 2>gpurun_out/bench11.err; cut -c1-300 gpurun_out/bench11.log
