mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/gputest27.log 2>&1; tail -12 gpurun_out/gputest27.log | cut -c1-400
python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench22.log 2>gpurun_out/bench22.err; cut -c1-330 gpurun_out/bench22.log; tail -3 gpurun_out/bench22.err
