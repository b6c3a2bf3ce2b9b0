mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/gputest16.log 2>&1; tail -15 gpurun_out/gputest16.log | cut -c1-250
python bench.py --steps 50 --warmup 10 > gpurun_out/bench12.log 2>gpurun_out/bench12.err; cut -c1-300 gpurun_out/bench12.log; tail -3 gpurun_out/bench12.err
