# r02am (2 GPUs): multi-context / multi-GPU tests (incl. the new voxel-shape test) and the bench at 2 GPUs at HEAD (spp and tiles, with the frame check)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_progressive_multictx.py -x -q -m gpu > gpurun_out/r02am_tests.log 2>&1; tail -4 gpurun_out/r02am_tests.log
for sh in spp tiles; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --shard $sh > gpurun_out/r02am_bench_g2_$sh.json 2> gpurun_out/r02am_bench_g2_$sh.err; cut -c1-160 gpurun_out/r02am_bench_g2_$sh.json; tail -2 gpurun_out/r02am_bench_g2_$sh.err
done
