# r02af: HEAD: full GPU suite, smoke, --set full at the bench configuration, launch list, the default bench line
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02af_tests_all.log 2>&1; tail -3 gpurun_out/r02af_tests_all.log
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/r02af_smoke.log 2>&1; tail -3 gpurun_out/r02af_smoke.log | cut -c1-160
CMD="python bench.py --steps 2 --warmup 3 --no-ref-cuda --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02af_launches_bench_c3.csv $CMD > gpurun_out/r02af_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_coop -s 1 -c 1 -o gpurun_out/r02af_coop_c3_64spp -f python profiles/profile_render.py C3 64 2 > gpurun_out/r02af_ncu_full.log 2>&1; tail -1 gpurun_out/r02af_ncu_full.log
timeout 900 python bench.py > gpurun_out/r02af_bench_c3.json 2> gpurun_out/r02af_bench_c3.err; cut -c1-200 gpurun_out/r02af_bench_c3.json; tail -2 gpurun_out/r02af_bench_c3.err
