mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/gputest21.log 2>&1; tail -12 gpurun_out/gputest21.log | cut -c1-250
python bench.py --steps 50 --warmup 10 > gpurun_out/bench17.log 2>gpurun_out/bench17.err; cut -c1-300 gpurun_out/bench17.log; tail -3 gpurun_out/bench17.err
GLIS_FUSE_TPRELU_BWD=0 python bench.py --steps 50 --warmup 10 > gpurun_out/bench17b.log 2>gpurun_out/bench17b.err; cut -c100-260 gpurun_out/bench17b.log
