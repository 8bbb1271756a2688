# r02d: new GPU tests (progressive, per-context camera, kernel id, coop vs lane), full-size diagnosis against the reference CUDA build
timeout 900 python -m pytest tests/test_gpu_progressive_multictx.py -x -q -m gpu > gpurun_out/r02d_tests_new.log 2>&1; tail -15 gpurun_out/r02d_tests_new.log
timeout 1200 python profiles/diag_full_frame.py 64 1,40 > gpurun_out/r02d_diag.log 2>&1; tail -12 gpurun_out/r02d_diag.log | cut -c1-900
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02d_tests_all.log 2>&1; tail -8 gpurun_out/r02d_tests_all.log
