mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/gputest32.log 2>&1; tail -5 gpurun_out/gputest32.log | cut -c1-300
python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench27.log 2>gpurun_out/bench27.err; cut -c1-330 gpurun_out/bench27.log; tail -3 gpurun_out/bench27.err
python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench27b.log 2>gpurun_out/bench27b.err; cut -c100-260 gpurun_out/bench27b.log
