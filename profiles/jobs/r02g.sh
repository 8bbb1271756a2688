# r02g: exact tie order (tie_key) in all kernels: variants agree, full GPU suite incl. the full-size sha test, bench line
timeout 300 python profiles/sweep_variants.py C3 8 1,40,41,11 > gpurun_out/r02g_ab_c3.log 2>&1; cat gpurun_out/r02g_ab_c3.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02g_tests_all.log 2>&1; tail -8 gpurun_out/r02g_tests_all.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-ref-cuda > gpurun_out/r02g_bench_c3.json 2> gpurun_out/r02g_bench_c3.err; cat gpurun_out/r02g_bench_c3.json | cut -c1-1500; tail -3 gpurun_out/r02g_bench_c3.err
