mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 700 --csv --log-file gpurun_out/launches_warm.csv python bench.py --steps 2 --warmup 3 --no-graph > gpurun_out/ncu_warm.log 2>&1; tail -1 gpurun_out/ncu_warm.log | cut -c1-100
