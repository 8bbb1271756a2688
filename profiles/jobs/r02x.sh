# r02x: owner-table lookup + ballots-first push: A/B against the r02u numbers (C3 8 spp: 17.7 ms; C5 2 spp: 47.4 ms), full GPU suite
timeout 300 python profiles/sweep_variants.py C3 8 0,1,11 > gpurun_out/r02x_ab_c3.log 2>&1; cat gpurun_out/r02x_ab_c3.log
timeout 300 python profiles/sweep_variants.py C5 2 0 > gpurun_out/r02x_ab_c5.log 2>&1; cat gpurun_out/r02x_ab_c5.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02x_tests_all.log 2>&1; tail -3 gpurun_out/r02x_tests_all.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02x_bench_c3.json 2> gpurun_out/r02x_bench_c3.err; cut -c1-200 gpurun_out/r02x_bench_c3.json; grep -o '"live": {[^}]*}' gpurun_out/r02x_bench_c3.json | cut -c1-300
