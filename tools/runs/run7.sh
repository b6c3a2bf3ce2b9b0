mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/gputest17.log 2>&1; tail -5 gpurun_out/gputest17.log | cut -c1-250
timeout 300 python tools/tc_microbench.py 1 > gpurun_out/tcmb17.log 2>&1; cat gpurun_out/tcmb17.log | tail -16
python bench.py --steps 50 --warmup 10 > gpurun_out/bench13.log 2>gpurun_out/bench13.err; cut -c1-300 gpurun_out/bench13.log; tail -3 gpurun_out/bench13.err
