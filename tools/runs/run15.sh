mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/gputest24.log 2>&1; tail -4 gpurun_out/gputest24.log | cut -c1-300
python bench.py --steps 100 --warmup 20 --kernel-table gpurun_out/kernels_r1n.json > gpurun_out/bench19.log 2>gpurun_out/bench19.err; cut -c1-330 gpurun_out/bench19.log; tail -3 gpurun_out/bench19.err
python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/plain19.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1n.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu19.log 2>&1; tail -1 gpurun_out/ncu19.log | cut -c1-200
