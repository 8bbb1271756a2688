# r02h: coop v4 (certain-hit bounds, exact evaluation deferred to full rings): A/B + parity subset
timeout 300 python profiles/sweep_variants.py C3 8 1,40,41,44 > gpurun_out/r02h_ab_c3.log 2>&1; cat gpurun_out/r02h_ab_c3.log
timeout 300 python profiles/sweep_variants.py C5 2 11,40 > gpurun_out/r02h_ab_c5.log 2>&1; cat gpurun_out/r02h_ab_c5.log
timeout 300 python profiles/sweep_variants.py C2 10 1,40 > gpurun_out/r02h_ab_c2.log 2>&1; cat gpurun_out/r02h_ab_c2.log
RT_RENDER_VARIANT=40 timeout 900 python -m pytest tests -x -q -m gpu -k "render_matches or closest_hit or flat_list or octree_and_flat or full_size or cooperative or dropin_octree" > gpurun_out/r02h_tests.log 2>&1; tail -5 gpurun_out/r02h_tests.log
