# r02o: full GPU suite at HEAD (new full-size C4 and 1 M-sphere tests), bench lines for C4 (USE_FP16, with its roofline) and C5
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02o_tests_all.log 2>&1; tail -6 gpurun_out/r02o_tests_all.log
timeout 900 python bench.py --config C4 --steps 2 --warmup 3 --no-ref-cuda > gpurun_out/r02o_bench_c4.json 2> gpurun_out/r02o_bench_c4.err; cut -c1-2600 gpurun_out/r02o_bench_c4.json; tail -2 gpurun_out/r02o_bench_c4.err
timeout 900 python bench.py --config C5 --steps 2 --warmup 3 --no-ref-cuda --no-cpu-baseline > gpurun_out/r02o_bench_c5.json 2> gpurun_out/r02o_bench_c5.err; cut -c1-2600 gpurun_out/r02o_bench_c5.json; tail -2 gpurun_out/r02o_bench_c5.err
