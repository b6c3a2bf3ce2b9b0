mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke3.log 2>&1; tail -2 gpurun_out/smoke3.log | cut -c1-250
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/gputest37.log 2>&1; tail -4 gpurun_out/gputest37.log | cut -c1-300
python bench.py --steps 100 --warmup 20 > gpurun_out/bench_final2.log 2>gpurun_out/bench_final2.err; cut -c1-330 gpurun_out/bench_final2.log; tail -2 gpurun_out/bench_final2.err
