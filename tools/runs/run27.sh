mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/gputest34.log 2>&1; tail -6 gpurun_out/gputest34.log | cut -c1-400
python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench29.log 2>gpurun_out/bench29.err; cut -c100-260 gpurun_out/bench29.log; tail -3 gpurun_out/bench29.err
GLIS_LIS_FUSED=0 python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench29b.log 2>gpurun_out/bench29b.err; cut -c100-260 gpurun_out/bench29b.log
python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench29c.log 2>gpurun_out/bench29c.err; cut -c100-260 gpurun_out/bench29c.log
