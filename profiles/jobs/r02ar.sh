# r02ar: per-voxel exit cap in the cooperative pre-filter: timing at C3 / C5, then the GPU suite
mkdir -p gpurun_out
python profiles/sweep_variants.py C3 8 0,1 > gpurun_out/r02ar_c3.log 2>&1; grep variant gpurun_out/r02ar_c3.log
python profiles/sweep_variants.py C5 2 0 > gpurun_out/r02ar_c5.log 2>&1; grep variant gpurun_out/r02ar_c5.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02ar_tests_all.log 2>&1; tail -3 gpurun_out/r02ar_tests_all.log
