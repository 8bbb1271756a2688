# r02ab (2 GPUs): sharded frame queued without a host round trip in the middle (stats read at the end): bench at 2 GPUs, NCCL tests
for sh in spp tiles; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --shard $sh > gpurun_out/r02ab_bench_g2_$sh.json 2> gpurun_out/r02ab_bench_g2_$sh.err
  cut -c1-220 gpurun_out/r02ab_bench_g2_$sh.json; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r02ab_bench_g2_$sh.json; grep -o '"identical_to_1gpu_frame": [a-z]*, "pixels_identical": [0-9.e-]*' gpurun_out/r02ab_bench_g2_$sh.json; grep -o '"phases_ms_rank0": {[^}]*}' gpurun_out/r02ab_bench_g2_$sh.json; tail -2 gpurun_out/r02ab_bench_g2_$sh.err | cut -c1-200
done
timeout 600 python -m pytest tests/test_gpu_progressive_multictx.py -x -q -m gpu -k "nccl or multi_gpu" > gpurun_out/r02ab_tests.log 2>&1; tail -3 gpurun_out/r02ab_tests.log
