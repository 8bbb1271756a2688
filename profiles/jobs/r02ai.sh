# r02ai: flat voxels in slab scenes (choose_grid): full GPU suite, smoke, bench C3 and C5 (no reference legs), --set full at the bench configuration
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02ai_tests_all.log 2>&1; tail -3 gpurun_out/r02ai_tests_all.log
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/r02ai_smoke.log 2>&1; tail -3 gpurun_out/r02ai_smoke.log | cut -c1-160
timeout 600 python bench.py --no-ref-cuda --no-cpu-baseline > gpurun_out/r02ai_bench_c3.json 2> gpurun_out/r02ai_bench_c3.err; cut -c1-220 gpurun_out/r02ai_bench_c3.json; tail -2 gpurun_out/r02ai_bench_c3.err
timeout 900 python bench.py --config C5 --steps 2 --warmup 3 --no-ref-cuda --no-cpu-baseline > gpurun_out/r02ai_bench_c5.json 2> gpurun_out/r02ai_bench_c5.err; cut -c1-220 gpurun_out/r02ai_bench_c5.json; tail -2 gpurun_out/r02ai_bench_c5.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_coop -s 1 -c 1 -o gpurun_out/r02ai_coop_c3_64spp -f python profiles/profile_render.py C3 64 2 > gpurun_out/r02ai_ncu_full.log 2>&1; tail -1 gpurun_out/r02ai_ncu_full.log
