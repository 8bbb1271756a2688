# r02l (2 GPUs): kernel-switch sweep; multi-GPU C-ABI test; bench at 2 GPUs (tiles and spp) with the frame check; CLI with RT_GPUS=2
timeout 300 python profiles/sweep_coop_threshold.py 3840 2160 8 > gpurun_out/r02l_sweep_4k.log 2>&1; cat gpurun_out/r02l_sweep_4k.log
timeout 300 python profiles/sweep_coop_threshold.py 1200 800 10 > gpurun_out/r02l_sweep_1200.log 2>&1; cat gpurun_out/r02l_sweep_1200.log
timeout 600 python -m pytest tests/test_gpu_progressive_multictx.py -x -q -m gpu > gpurun_out/r02l_tests.log 2>&1; tail -5 gpurun_out/r02l_tests.log
for sh in tiles spp; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --shard $sh > gpurun_out/r02l_bench_g2_$sh.json 2> gpurun_out/r02l_bench_g2_$sh.err; cut -c1-2500 gpurun_out/r02l_bench_g2_$sh.json; tail -3 gpurun_out/r02l_bench_g2_$sh.err
done
cd dd2360-raytracing_b200
RT_NUM_SPHERES=100000 RT_SPHERES_PER_LEAF=300 RT_NX=3840 RT_NY=2160 RT_NS=16 RT_VERBOSE=1 ./RayTracing 3 2> ../gpurun_out/r02l_cli_g1.err; sha256sum output.ppm > ../gpurun_out/r02l_cli_sha.txt
RT_GPUS=2 RT_NUM_SPHERES=100000 RT_SPHERES_PER_LEAF=300 RT_NX=3840 RT_NY=2160 RT_NS=16 RT_VERBOSE=1 ./RayTracing 3 2> ../gpurun_out/r02l_cli_g2.err; sha256sum output.ppm >> ../gpurun_out/r02l_cli_sha.txt
RT_GPUS=2 RT_SHARD=spp RT_NUM_SPHERES=100000 RT_SPHERES_PER_LEAF=300 RT_NX=3840 RT_NY=2160 RT_NS=16 RT_VERBOSE=1 ./RayTracing 1 2> ../gpurun_out/r02l_cli_g2_spp.err
cd ..; cat gpurun_out/r02l_cli_sha.txt; tail -2 gpurun_out/r02l_cli_g1.err gpurun_out/r02l_cli_g2.err gpurun_out/r02l_cli_g2_spp.err
