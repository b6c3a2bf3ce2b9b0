# round 2, session 2, call 6: what the fused epilogue (pre-activation + bf16 planes: 2-byte stores) costs the conv kernels
mkdir -p gpurun_out
echo "plain fp32 output: halo/one-box kernel as planned, then halo with no stores (t4), one-box (h0), one-box no stores (d1)" > gpurun_out/s2_epi.log
python tools/tc_microbench.py k32 t4 h0 d1 >> gpurun_out/s2_epi.log 2>&1
echo "fused epilogue (bias + TPReLU, preact + hi/lo planes)" >> gpurun_out/s2_epi.log
TCMB_EPILOGUE=1 python tools/tc_microbench.py k32 t4 h0 d1 >> gpurun_out/s2_epi.log 2>&1
cat gpurun_out/s2_epi.log
