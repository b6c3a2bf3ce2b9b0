# round 2, session 2, call 1: full GPU suite, bench (default), launch list, ncu full of the tcgen05 kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; tail -3 gpurun_out/s2_pytest.log
timeout 300 python bench.py > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; cat gpurun_out/s2_bench.json | cut -c1-1500
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/s2_b_nograph.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1200 --csv --log-file gpurun_out/s2_launches_warm.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/s2_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_conv|tc_wgrad' --launch-skip 120 -c 40 -o gpurun_out/s2_tc_full -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/s2_ncu2.log 2>&1
ls -la gpurun_out | tail
