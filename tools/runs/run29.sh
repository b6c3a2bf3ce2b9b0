mkdir -p gpurun_out
: > gpurun_out/bisect.log
for cfg in "A=1" "GLIS_LIS_FUSED=0" "GLIS_TC_PM=0" "GLIS_SPLIT_K_FORWARD=0" "GLIS_OVERLAP_WGRAD=0"; do
  fails=0
  for i in 1 2 3 4 5 6; do
    env $cfg timeout 120 python -m pytest tests/test_gpu_parity.py -q -x -k "config4 and bf16x3" > gpurun_out/bisect_one.log 2>&1 || { fails=$((fails+1)); grep -m1 "AssertionError: (" gpurun_out/bisect_one.log | cut -c1-160 >> gpurun_out/bisect.log; }
  done
  echo "$cfg fails=$fails/6" >> gpurun_out/bisect.log
done
cat gpurun_out/bisect.log
