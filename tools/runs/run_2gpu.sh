mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 20 > gpurun_out/bench_2gpu_$tag.log 2>gpurun_out/bench_2gpu_$tag.err; echo "$tag $(tail -1 gpurun_out/bench_2gpu_$tag.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4))')"; }
run base A=1
run cta4 NCCL_MAX_CTAS=4
run cta8 NCCL_MAX_CTAS=8
run cta16 NCCL_MAX_CTAS=16
