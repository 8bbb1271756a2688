# r02ak: evidence at HEAD (flat voxels): full GPU suite, smoke, launch list of the bench command, --set full of the C5 kernel,
# the default bench line exactly as the driver runs it (live reference-CUDA leg, CPU baseline), the reference arm, the C5 line
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02ak_tests_all.log 2>&1; tail -3 gpurun_out/r02ak_tests_all.log
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/r02ak_smoke.log 2>&1; tail -3 gpurun_out/r02ak_smoke.log | cut -c1-160
CMD="python bench.py --steps 2 --warmup 3 --no-ref-cuda --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02ak_launches_bench_c3.csv $CMD > gpurun_out/r02ak_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_coop -s 1 -c 1 -o gpurun_out/r02ak_coop_c5_2spp -f python profiles/profile_render.py C5 2 2 > gpurun_out/r02ak_ncu_full_c5.log 2>&1; tail -1 gpurun_out/r02ak_ncu_full_c5.log
timeout 900 python bench.py > gpurun_out/r02ak_bench_c3.json 2> gpurun_out/r02ak_bench_c3.err; cut -c1-200 gpurun_out/r02ak_bench_c3.json; tail -2 gpurun_out/r02ak_bench_c3.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02ak_bench_ref_arm.json 2> gpurun_out/r02ak_bench_ref_arm.err; cut -c1-300 gpurun_out/r02ak_bench_ref_arm.json
timeout 900 python bench.py --config C5 --steps 2 --warmup 3 --no-ref-cuda > gpurun_out/r02ak_bench_c5.json 2> gpurun_out/r02ak_bench_c5.err; cut -c1-200 gpurun_out/r02ak_bench_c5.json
