# r02w (8 GPUs): final multi-GPU lines at HEAD: C3 spp, C5 spp, C3 at 4 GPUs; the torch.distributed NCCL frame test
run() { # n, tag, args...
  local n=$1 tag=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n "$@" > gpurun_out/r02w_bench_$tag.json 2> gpurun_out/r02w_bench_$tag.err
  cut -c1-230 gpurun_out/r02w_bench_$tag.json; grep -o '"e2e": {[^}]*}' gpurun_out/r02w_bench_$tag.json | cut -c1-200; grep -o '"phases_ms_rank0": {[^}]*}' gpurun_out/r02w_bench_$tag.json
}
run 8 g8_c3_spp --steps 5 --warmup 3 --shard spp
run 4 g4_c3_spp --steps 5 --warmup 3 --shard spp
run 8 g8_c5_spp --config C5 --steps 2 --warmup 3 --shard spp
timeout 600 python -m pytest tests/test_gpu_progressive_multictx.py -x -q -m gpu -k "nccl or multi_gpu" > gpurun_out/r02w_tests.log 2>&1; tail -3 gpurun_out/r02w_tests.log
