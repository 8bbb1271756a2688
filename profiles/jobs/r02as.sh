# r02as: HEAD (exit cap): --set full at the bench configuration, bench lines C3 and C5 (no reference legs)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_coop -s 1 -c 1 -o gpurun_out/r02as_coop_c3_64spp -f python profiles/profile_render.py C3 64 2 > gpurun_out/r02as_ncu_full.log 2>&1; tail -1 gpurun_out/r02as_ncu_full.log
timeout 600 python bench.py --no-ref-cuda --no-cpu-baseline > gpurun_out/r02as_bench_c3.json 2> gpurun_out/r02as_bench_c3.err; cut -c1-200 gpurun_out/r02as_bench_c3.json
timeout 900 python bench.py --config C5 --steps 2 --warmup 3 --no-ref-cuda --no-cpu-baseline > gpurun_out/r02as_bench_c5.json 2> gpurun_out/r02as_bench_c5.err; cut -c1-200 gpurun_out/r02as_bench_c5.json
