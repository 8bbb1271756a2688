# r02an (4 GPUs): the bench at 4 GPUs at HEAD (spp), with the frame check
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 3 --warmup 3 --shard spp > gpurun_out/r02an_bench_g4_spp.json 2> gpurun_out/r02an_bench_g4_spp.err; cut -c1-160 gpurun_out/r02an_bench_g4_spp.json; tail -2 gpurun_out/r02an_bench_g4_spp.err
