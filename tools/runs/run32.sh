mkdir -p gpurun_out
timeout 200 python tools/smoke_seeds.py > gpurun_out/smoke_seeds.log 2>&1; tail -9 gpurun_out/smoke_seeds.log | cut -c1-200
