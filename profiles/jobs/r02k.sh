# r02k: coop A/B: 40 = v4 form, 46 = + pipelined loads, 47 = + atomic push, 48 = pipelined at 5 blocks/SM
timeout 300 python profiles/sweep_variants.py C3 8 40,46,47,48,41 > gpurun_out/r02k_ab_c3.log 2>&1; cat gpurun_out/r02k_ab_c3.log
timeout 300 python profiles/sweep_variants.py C5 2 40,46,47 > gpurun_out/r02k_ab_c5.log 2>&1; cat gpurun_out/r02k_ab_c5.log
