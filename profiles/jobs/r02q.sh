# r02q (8 GPUs): bench at 8 GPUs (C3 spp and tiles, C5 spp) with the frame check against the 1-GPU frame; CLI with RT_GPUS=8
run() { # tag, args...
  local tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 "$@" > gpurun_out/r02q_bench_g8_$tag.json 2> gpurun_out/r02q_bench_g8_$tag.err
  cut -c1-260 gpurun_out/r02q_bench_g8_$tag.json; grep -o '"e2e": {[^}]*}' gpurun_out/r02q_bench_g8_$tag.json | cut -c1-200; grep -o '"frame_check": {[^}]*}' gpurun_out/r02q_bench_g8_$tag.json | cut -c1-400
}
run c3_spp --steps 5 --warmup 3 --shard spp
run c3_tiles --steps 5 --warmup 3 --shard tiles
run c5_spp --config C5 --steps 2 --warmup 3 --shard spp
cd dd2360-raytracing_b200
RT_GPUS=8 RT_NUM_SPHERES=100000 RT_SPHERES_PER_LEAF=300 RT_NX=3840 RT_NY=2160 RT_NS=64 RT_VERBOSE=1 ./RayTracing 3 2> ../gpurun_out/r02q_cli_g8.err; sha256sum output.ppm > ../gpurun_out/r02q_cli_sha.txt
RT_GPUS=1 RT_NUM_SPHERES=100000 RT_SPHERES_PER_LEAF=300 RT_NX=3840 RT_NY=2160 RT_NS=64 RT_VERBOSE=1 ./RayTracing 3 2> ../gpurun_out/r02q_cli_g1.err; sha256sum output.ppm >> ../gpurun_out/r02q_cli_sha.txt
cd ..; cat gpurun_out/r02q_cli_sha.txt; tail -n 2 gpurun_out/r02q_cli_g8.err; tail -n 2 gpurun_out/r02q_cli_g1.err
