mkdir -p gpurun_out
timeout 300 python tools/debug_overlap.py bf16x3 > gpurun_out/overlap14.log 2>&1; cat gpurun_out/overlap14.log | cut -c1-260
