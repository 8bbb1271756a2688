# r02r: build timeline; shared-memory parking variants of the cooperative kernel
timeout 300 python profiles/time_build.py C3 > gpurun_out/r02r_time_build.log 2>&1; cat gpurun_out/r02r_time_build.log
timeout 300 python profiles/sweep_variants.py C3 8 40,47,48,49,41 > gpurun_out/r02r_ab_c3.log 2>&1; cat gpurun_out/r02r_ab_c3.log
timeout 300 python profiles/sweep_variants.py C5 2 45,47 > gpurun_out/r02r_ab_c5.log 2>&1; cat gpurun_out/r02r_ab_c5.log
