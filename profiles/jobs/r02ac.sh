# r02ac: per-lane candidate queues (53 / 54) against the ring form (55 / 56), same shared-memory footprint
timeout 300 python profiles/sweep_variants.py C3 8 55,53,1 > gpurun_out/r02ac_ab_c3.log 2>&1; cat gpurun_out/r02ac_ab_c3.log
timeout 300 python profiles/sweep_variants.py C5 2 56,54 > gpurun_out/r02ac_ab_c5.log 2>&1; cat gpurun_out/r02ac_ab_c5.log
RT_RENDER_VARIANT=53 timeout 900 python -m pytest tests -x -q -m gpu -k "render_matches or closest_hit or flat_list or octree_and_flat or cooperative or full_size_frame" > gpurun_out/r02ac_tests.log 2>&1; tail -3 gpurun_out/r02ac_tests.log
