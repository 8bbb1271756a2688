# r02m: evidence for the bench kernel at HEAD: bench line, ncu launch list of the same command, --set full at the bench config (64 spp)
CMD="python bench.py --steps 2 --warmup 3 --no-ref-cuda --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/r02m_bench_plain.json 2> gpurun_out/r02m_bench_plain.err; cut -c1-300 gpurun_out/r02m_bench_plain.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02m_launches_bench_c3.csv $CMD > gpurun_out/r02m_ncu_launches.log 2>&1; tail -2 gpurun_out/r02m_ncu_launches.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_coop -s 1 -c 1 -o gpurun_out/r02m_coop_c3_64spp -f python profiles/profile_render.py C3 64 2 > gpurun_out/r02m_ncu_full.log 2>&1; tail -2 gpurun_out/r02m_ncu_full.log
