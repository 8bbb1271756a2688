# r02i: ncu --set full of k_render_coop (v4) at C3 / 8 spp and at C5 / 2 spp
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_coop -s 1 -c 1 -o gpurun_out/r02i_coop_c3_8spp -f python profiles/profile_render.py C3 8 2 > gpurun_out/r02i_ncu_c3.log 2>&1; tail -2 gpurun_out/r02i_ncu_c3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_coop -s 1 -c 1 -o gpurun_out/r02i_coop_c5_2spp -f python profiles/profile_render.py C5 2 2 > gpurun_out/r02i_ncu_c5.log 2>&1; tail -2 gpurun_out/r02i_ncu_c5.log
