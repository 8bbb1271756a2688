# r02s: new defaults (7 blocks/SM + parked pixel state; scene upload through pinned staging): full GPU suite, A/B, bench lines
timeout 300 python profiles/sweep_variants.py C5 2 45,50,51,0 > gpurun_out/r02s_ab_c5.log 2>&1; cat gpurun_out/r02s_ab_c5.log
timeout 300 python profiles/sweep_variants.py C3 8 40,48,0 > gpurun_out/r02s_ab_c3.log 2>&1; cat gpurun_out/r02s_ab_c3.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02s_tests_all.log 2>&1; tail -4 gpurun_out/r02s_tests_all.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-ref-cuda --no-cpu-baseline > gpurun_out/r02s_bench_c3.json 2> gpurun_out/r02s_bench_c3.err; cut -c1-200 gpurun_out/r02s_bench_c3.json; grep -o '"e2e": {.*"includes' gpurun_out/r02s_bench_c3.json | cut -c1-600
timeout 300 python profiles/time_build.py C3 2>&1 | grep upload | tail -1
