# r02p: the bench line as the driver runs it (N = 1, defaults), including the live run of the reference's CUDA build, and the reference arm
timeout 900 python bench.py > gpurun_out/r02p_bench_c3.json 2> gpurun_out/r02p_bench_c3.err; cut -c1-4000 gpurun_out/r02p_bench_c3.json; tail -3 gpurun_out/r02p_bench_c3.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02p_bench_reference_arm.json 2> gpurun_out/r02p_bench_reference_arm.err; cut -c1-600 gpurun_out/r02p_bench_reference_arm.json
