# r02u: 8 blocks/SM as the default for two candidates per lane; four candidates per lane at 8 blocks (C5)
timeout 300 python profiles/sweep_variants.py C5 2 51,52 > gpurun_out/r02u_ab_c5.log 2>&1; cat gpurun_out/r02u_ab_c5.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02u_tests_all.log 2>&1; tail -3 gpurun_out/r02u_tests_all.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-ref-cuda --no-cpu-baseline > gpurun_out/r02u_bench_c3.json 2> gpurun_out/r02u_bench_c3.err; cut -c1-200 gpurun_out/r02u_bench_c3.json
