mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log | cut -c1-300
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/gputest35.log 2>&1; tail -4 gpurun_out/gputest35.log | cut -c1-300
python bench.py > gpurun_out/bench_final.log 2>gpurun_out/bench_final.err; cut -c1-330 gpurun_out/bench_final.log; tail -2 gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref.log; tail -2 gpurun_out/bench_ref.err
python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/plain30.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_r1q.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu30.log 2>&1; tail -1 gpurun_out/ncu30.log | cut -c1-100
ncu --set full --clock-control none --import-source on -k 'regex:tc_pm_kernel|lis_chain_kernel|tprelu_fwd_planes' -s 30 -c 12 -f -o gpurun_out/r1q_pm python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu30b.log 2>&1; tail -1 gpurun_out/ncu30b.log | cut -c1-100
