mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/gputest26.log 2>&1; tail -6 gpurun_out/gputest26.log | cut -c1-300
python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench21.log 2>gpurun_out/bench21.err; cut -c1-330 gpurun_out/bench21.log; tail -3 gpurun_out/bench21.err
GLIS_SPLIT_K_FORWARD=0 python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench21b.log 2>gpurun_out/bench21b.err; cut -c100-260 gpurun_out/bench21b.log
