# r02n: four candidates per lane (variants 45/46) against two (40): C3, C5
timeout 300 python profiles/sweep_variants.py C3 8 40,45,46 > gpurun_out/r02n_ab_c3.log 2>&1; cat gpurun_out/r02n_ab_c3.log
timeout 300 python profiles/sweep_variants.py C5 2 40,45,46 > gpurun_out/r02n_ab_c5.log 2>&1; cat gpurun_out/r02n_ab_c5.log
RT_RENDER_VARIANT=45 timeout 900 python -m pytest tests -x -q -m gpu -k "render_matches or closest_hit or flat_list or octree_and_flat or cooperative" > gpurun_out/r02n_tests.log 2>&1; tail -3 gpurun_out/r02n_tests.log
