# r02e: coop v2 (pairs, MATCH/REDUX drain) A/B; tie rule + UB value: the full-size frame against the reference CUDA frame; new tests
timeout 300 python profiles/sweep_variants.py C3 8 1,40,44,41,42,11 > gpurun_out/r02e_ab_c3.log 2>&1; cat gpurun_out/r02e_ab_c3.log
timeout 300 python profiles/sweep_variants.py C2 10 1,40 > gpurun_out/r02e_ab_c2.log 2>&1; cat gpurun_out/r02e_ab_c2.log
timeout 300 python profiles/sweep_variants.py C5 2 11,40,41 > gpurun_out/r02e_ab_c5.log 2>&1; cat gpurun_out/r02e_ab_c5.log
timeout 600 python profiles/diag_full_frame.py 64 0,1,11 > gpurun_out/r02e_diag.log 2>&1; grep -E "^variant|part from" gpurun_out/r02e_diag.log | cut -c1-700
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02e_tests_all.log 2>&1; tail -8 gpurun_out/r02e_tests_all.log
