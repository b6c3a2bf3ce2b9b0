# usage: bash tools/dp_bench.sh N [ENV=VAL ...]   -> ms per step at N GPUs
N=$1; shift
env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('%d GPUs %s: %.4f ms  %.0f img/s' % (d['n_gpus'], '$*', d['ms_per_step'], d['value']))"
