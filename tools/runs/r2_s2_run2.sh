# round 2, session 2, call 2: launch list (warm) + ncu full of the tcgen05 kernels, exported on the box (the .ncu-rep stays there)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/s2_b_nograph.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1200 --csv --log-file gpurun_out/s2_launches_warm.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/s2_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_conv|tc_wgrad' --launch-skip 96 -c 32 -o /tmp/s2_tc_full -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/s2_ncu2.log 2>&1
ncu -i /tmp/s2_tc_full.ncu-rep --page raw --csv > /tmp/s2_tc_full_raw.csv 2>/dev/null
python tools/ncu_extract.py < /tmp/s2_tc_full_raw.csv > gpurun_out/s2_tc_ncu.csv
ls -la gpurun_out /tmp/s2_tc_full* | tail
