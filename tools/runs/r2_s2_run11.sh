# round 2, session 2, call 11: fresh warm launch list
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1200 --csv --log-file gpurun_out/s2_launches_warm3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/s2_ncu4.log 2>&1
tail -2 gpurun_out/s2_ncu4.log
