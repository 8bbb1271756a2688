# r02t: more parked state (21 words + the ray re-read from the warp's ray table): 6 / 7 / 8 blocks per SM
timeout 300 python profiles/sweep_variants.py C3 8 47,48,49,0 > gpurun_out/r02t_ab_c3.log 2>&1; cat gpurun_out/r02t_ab_c3.log
timeout 300 python profiles/sweep_variants.py C5 2 50,51,0 > gpurun_out/r02t_ab_c5.log 2>&1; cat gpurun_out/r02t_ab_c5.log
RT_RENDER_VARIANT=49 timeout 900 python -m pytest tests -x -q -m gpu -k "render_matches or closest_hit or flat_list or octree_and_flat or cooperative or progressive" > gpurun_out/r02t_tests.log 2>&1; tail -3 gpurun_out/r02t_tests.log
