mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/gputest33.log 2>&1; tail -12 gpurun_out/gputest33.log | cut -c1-400
python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench28.log 2>gpurun_out/bench28.err; cut -c1-330 gpurun_out/bench28.log; tail -3 gpurun_out/bench28.err
GLIS_LIS_FUSED=0 python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench28b.log 2>gpurun_out/bench28b.err; cut -c100-260 gpurun_out/bench28b.log
python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench28c.log 2>gpurun_out/bench28c.err; cut -c100-260 gpurun_out/bench28c.log
