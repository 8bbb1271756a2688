# r02b: first run of the warp-cooperative kernel (variant 40): parity subset, then A/B against k_render (1) and the pool (11)
RT_RENDER_VARIANT=40 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "render_matches or closest_hit or flat_list or octree_and_flat or nan_pixel or upstream_seeding_matches" > gpurun_out/r02b_tests.log 2>&1
tail -15 gpurun_out/r02b_tests.log
timeout 300 python profiles/sweep_variants.py C3 8 1,40,41,42,43,11 > gpurun_out/r02b_ab_c3.log 2>&1; cat gpurun_out/r02b_ab_c3.log
timeout 300 python profiles/sweep_variants.py C2 10 1,40 > gpurun_out/r02b_ab_c2.log 2>&1; cat gpurun_out/r02b_ab_c2.log
timeout 300 python profiles/sweep_variants.py C5 2 11,40,1 > gpurun_out/r02b_ab_c5.log 2>&1; cat gpurun_out/r02b_ab_c5.log
