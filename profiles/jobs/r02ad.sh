# r02ad: warp-level discriminant cull of the big-sphere (prolog) tests
timeout 300 python profiles/sweep_variants.py C3 8 0,1 > gpurun_out/r02ad_ab_c3.log 2>&1; cat gpurun_out/r02ad_ab_c3.log
timeout 300 python profiles/sweep_variants.py C5 2 0 > gpurun_out/r02ad_ab_c5.log 2>&1; cat gpurun_out/r02ad_ab_c5.log
timeout 900 python -m pytest tests -x -q -m gpu -k "render_matches or closest_hit or flat_list or octree_and_flat or cooperative or full_size_frame or one_million" > gpurun_out/r02ad_tests.log 2>&1; tail -3 gpurun_out/r02ad_tests.log
