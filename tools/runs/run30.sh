mkdir -p gpurun_out
timeout 400 python tools/flaky_config4.py 40 > gpurun_out/flaky.log 2>&1; tail -8 gpurun_out/flaky.log | cut -c1-220
