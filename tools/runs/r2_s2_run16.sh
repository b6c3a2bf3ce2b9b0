# round 2, session 2, call 16: every BASELINE config at one GPU; final warm launch list and ncu full of the tcgen05 kernels
mkdir -p gpurun_out
for c in 1 3 4 5a 5b; do
  timeout 600 python bench.py --config $c --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('config $c: %.4f ms/step  %.0f %s  e2e %.4f ms  launches %s' % (d['ms_per_step'], d['value'], d['unit'], d['e2e']['ms_per_step'], d['details'].get('launches_per_iteration')))" | tee -a gpurun_out/s2_configs.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1200 --csv --log-file gpurun_out/s2_launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/s2_ncu5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_conv|tc_wgrad' --launch-skip 96 -c 32 -o /tmp/s2_tc_final -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/s2_ncu6.log 2>&1
ncu -i /tmp/s2_tc_final.ncu-rep --page raw --csv > /tmp/s2_tc_final_raw.csv 2>/dev/null
python tools/ncu_extract.py < /tmp/s2_tc_final_raw.csv > gpurun_out/s2_tc_final_ncu.csv
timeout 600 ncu --set full --clock-control none -k regex:'wn_pack_wide|wn_project_kernel|linear_wgrad_project|tprelu_bwd|lis_chain|head_' --launch-skip 60 -c 40 -o /tmp/s2_pw_final -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/s2_ncu7.log 2>&1
ncu -i /tmp/s2_pw_final.ncu-rep --page raw --csv > /tmp/s2_pw_final_raw.csv 2>/dev/null
python tools/ncu_extract.py < /tmp/s2_pw_final_raw.csv > gpurun_out/s2_pw_final_ncu.csv
ls -la gpurun_out | tail -5
