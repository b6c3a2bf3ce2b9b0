mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/gputest20.log 2>&1; tail -3 gpurun_out/gputest20.log | cut -c1-250
python bench.py --steps 50 --warmup 10 > gpurun_out/bench16.log 2>gpurun_out/bench16.err; cut -c1-300 gpurun_out/bench16.log; tail -3 gpurun_out/bench16.err
GLIS_OVERLAP_WGRAD=0 python bench.py --steps 50 --warmup 10 > gpurun_out/bench16b.log 2>gpurun_out/bench16b.err; cut -c1-300 gpurun_out/bench16b.log | cut -c100-260
