mkdir -p gpurun_out
echo "== tiled" > gpurun_out/packmb.log; timeout 120 python tools/pack_microbench.py >> gpurun_out/packmb.log 2>&1
echo "== element-wise" >> gpurun_out/packmb.log; GLIS_PACK_TILED=0 timeout 120 python tools/pack_microbench.py >> gpurun_out/packmb.log 2>&1
cat gpurun_out/packmb.log | cut -c1-160
