# round 2, session 2, call 12: cta_group::2 pair kernel — per-layer time and agreement with the one-CTA kernels
mkdir -p gpurun_out
timeout 120 python tools/tc_microbench.py p0 p1 > gpurun_out/s2_pair_plain.log 2>&1; cat gpurun_out/s2_pair_plain.log
TCMB_EPILOGUE=1 timeout 120 python tools/tc_microbench.py p0 p1 > gpurun_out/s2_pair_epi.log 2>&1; cat gpurun_out/s2_pair_epi.log
