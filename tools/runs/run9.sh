mkdir -p gpurun_out
timeout 300 python tools/tc_microbench.py k1 k2 k4 k8 > gpurun_out/tcmb19.log 2>&1; cat gpurun_out/tcmb19.log | tail -16
python tools/edge_microbench.py > gpurun_out/edge19.log 2>&1; grep -i "head dgrad" gpurun_out/edge19.log
python -m pytest tests -m gpu -q -x > gpurun_out/gputest19.log 2>&1; tail -3 gpurun_out/gputest19.log | cut -c1-250
python bench.py --steps 50 --warmup 10 > gpurun_out/bench15.log 2>gpurun_out/bench15.err; cut -c1-300 gpurun_out/bench15.log; tail -3 gpurun_out/bench15.err
