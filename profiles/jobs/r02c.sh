# r02c: ncu --set full of the cooperative kernel (variant 40) at C3 / 8 spp, then the goldens job (r02a)
RT_VARIANT=40 timeout 300 python profiles/profile_render.py C3 8 2 counters > gpurun_out/r02c_plain.log 2>&1; cat gpurun_out/r02c_plain.log
RT_VARIANT=40 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_coop -s 1 -c 1 -o gpurun_out/r02c_coop_c3_8spp -f python profiles/profile_render.py C3 8 2 > gpurun_out/r02c_ncu.log 2>&1; tail -3 gpurun_out/r02c_ncu.log
bash profiles/jobs/r02a.sh
