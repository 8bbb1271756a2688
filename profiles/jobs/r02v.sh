# r02v: final evidence for the bench kernel at HEAD (k_render_coop<8,2,park>): ncu launch list of the bench command, --set full at the bench
# configuration, then the bench exactly as the driver runs it (live reference-CUDA leg, CPU baseline) and the reference arm
CMD="python bench.py --steps 2 --warmup 3 --no-ref-cuda --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/r02v_bench_plain.json 2> gpurun_out/r02v_bench_plain.err; cut -c1-200 gpurun_out/r02v_bench_plain.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02v_launches_bench_c3.csv $CMD > gpurun_out/r02v_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_coop -s 1 -c 1 -o gpurun_out/r02v_coop_c3_64spp -f python profiles/profile_render.py C3 64 2 > gpurun_out/r02v_ncu_full.log 2>&1; tail -1 gpurun_out/r02v_ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_coop -s 1 -c 1 -o gpurun_out/r02v_coop_c5_2spp -f python profiles/profile_render.py C5 2 2 > gpurun_out/r02v_ncu_full_c5.log 2>&1; tail -1 gpurun_out/r02v_ncu_full_c5.log
timeout 900 python bench.py > gpurun_out/r02v_bench_c3.json 2> gpurun_out/r02v_bench_c3.err; cut -c1-200 gpurun_out/r02v_bench_c3.json; tail -2 gpurun_out/r02v_bench_c3.err
