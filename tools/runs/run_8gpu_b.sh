mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 60 --warmup 10 > gpurun_out/bench_8gpu_$tag.log 2>gpurun_out/bench_8gpu_$tag.err; echo "$tag $(tail -1 gpurun_out/bench_8gpu_$tag.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4))')"; }
run b2 GLIS_DP_BUCKET_MB=2
run b2cta8 GLIS_DP_BUCKET_MB=2 NCCL_MAX_CTAS=8
run b2cta16 GLIS_DP_BUCKET_MB=2 NCCL_MAX_CTAS=16
