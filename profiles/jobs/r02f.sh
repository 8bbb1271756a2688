# r02f: coop v3 (two native shared atomics in the drain) A/B, then the pixel-by-pixel diff against the reference CUDA frame
timeout 300 python profiles/sweep_variants.py C3 8 1,40,44,41 > gpurun_out/r02f_ab_c3.log 2>&1; cat gpurun_out/r02f_ab_c3.log
timeout 300 python profiles/sweep_variants.py C2 10 1,40 > gpurun_out/r02f_ab_c2.log 2>&1; cat gpurun_out/r02f_ab_c2.log
timeout 300 python profiles/sweep_variants.py C5 2 11,40 > gpurun_out/r02f_ab_c5.log 2>&1; cat gpurun_out/r02f_ab_c5.log
timeout 1200 python profiles/diff_full_frame.py > gpurun_out/r02f_diff.log 2>&1; tail -40 gpurun_out/r02f_diff.log | cut -c1-300
