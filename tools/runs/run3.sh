mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/gputest13.log 2>&1; tail -4 gpurun_out/gputest13.log
for s in 5 6 7 8; do python tools/debug_flip.py fp32 $s summary; done > gpurun_out/flip13.log 2>&1; grep -c seed gpurun_out/flip13.log; awk '{print $NF}' gpurun_out/flip13.log | sort -g | tail -5
python tools/edge_microbench.py > gpurun_out/edge13.log 2>&1; cat gpurun_out/edge13.log
python bench.py --steps 50 --warmup 10 > gpurun_out/bench10.log 2>gpurun_out/bench10.err; cut -c1-300 gpurun_out/bench10.log
