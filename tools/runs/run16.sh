mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/plain20.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 93 -c 3 -f -o gpurun_out/r1n_tc_conv $CMD > gpurun_out/ncu20a.log 2>&1
tail -2 gpurun_out/ncu20a.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k 'regex:tc_wgrad_kernel|tprelu_bwd_nhwc|rmsprop_kernel' -s 66 -c 22 -f -o gpurun_out/r1n_side $CMD > gpurun_out/ncu20b.log 2>&1
tail -2 gpurun_out/ncu20b.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
