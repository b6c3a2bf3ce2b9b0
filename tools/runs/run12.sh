mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/gputest22.log 2>&1; tail -8 gpurun_out/gputest22.log | cut -c1-250
python bench.py --steps 50 --warmup 10 > gpurun_out/bench18.log 2>gpurun_out/bench18.err; cut -c1-300 gpurun_out/bench18.log; tail -3 gpurun_out/bench18.err
