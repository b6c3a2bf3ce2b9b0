run() { env "$@" python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4f ms  %d launches' % (d['ms_per_step'], d['details']['launches_per_iteration']))"; }
echo "all on          : $(run X=1)"
echo "HALO=0          : $(run GLIS_TC_HALO=0)"
echo "MULTI_PREPARE=0 : $(run GLIS_MULTI_PREPARE=0)"
echo "MULTI_PROJECT=0 : $(run GLIS_MULTI_PROJECT=0)"
echo "FUSED_LIN=0     : $(run GLIS_FUSED_LINEAR_WGRAD=0)"
echo "all off         : $(run GLIS_TC_HALO=0 GLIS_MULTI_PREPARE=0 GLIS_MULTI_PROJECT=0 GLIS_FUSED_LINEAR_WGRAD=0)"
