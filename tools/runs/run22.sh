mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/gputest30.log 2>&1; tail -5 gpurun_out/gputest30.log | cut -c1-300
python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/bench25.log 2>gpurun_out/bench25.err; cut -c1-330 gpurun_out/bench25.log; tail -3 gpurun_out/bench25.err
python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/plain25.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1200 --csv --log-file gpurun_out/launches_r1p_warm.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu25.log 2>&1; tail -1 gpurun_out/ncu25.log | cut -c1-120
