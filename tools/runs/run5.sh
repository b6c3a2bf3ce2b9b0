mkdir -p gpurun_out
timeout 300 python tools/tc_microbench.py 1 2 4 > gpurun_out/tcmb15.log 2>&1; cat gpurun_out/tcmb15.log | tail -20
python tools/edge_microbench.py > gpurun_out/edge15.log 2>&1; grep -i "linear wgrad\|LIS wgrad" gpurun_out/edge15.log
