# r02a: full-size goldens from the reference's CUDA build (+ its timing with clocks), HEAD bench, work counters
bash tests/golden/gen_ref_cuda_full.sh all > gpurun_out/golden_full.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench_c3.json 2> gpurun_out/r02a_bench_c3.err
python profiles/profile_render.py C3 8 2 counters > gpurun_out/r02a_counters.log 2>&1
tail -3 gpurun_out/golden_full.log; cat gpurun_out/r02a_bench_c3.json; cat gpurun_out/r02a_counters.log
