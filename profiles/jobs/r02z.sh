# r02z: C5 bench line at HEAD (1 GPU); C2 and C1 lines for the table
timeout 900 python bench.py --config C5 --steps 2 --warmup 3 --no-ref-cuda --no-cpu-baseline > gpurun_out/r02z_bench_c5.json 2> gpurun_out/r02z_bench_c5.err; cut -c1-200 gpurun_out/r02z_bench_c5.json; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r02z_bench_c5.json
timeout 300 python bench.py --config C2 --steps 20 --warmup 5 --no-ref-cuda --no-cpu-baseline > gpurun_out/r02z_bench_c2.json 2> gpurun_out/r02z_bench_c2.err; cut -c1-200 gpurun_out/r02z_bench_c2.json
timeout 300 python bench.py --config C1 --steps 20 --warmup 5 --no-ref-cuda --no-cpu-baseline > gpurun_out/r02z_bench_c1.json 2> gpurun_out/r02z_bench_c1.err; cut -c1-200 gpurun_out/r02z_bench_c1.json
