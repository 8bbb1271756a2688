# r02al: A/B list prefetch at the voxel step (variants 53 / 54 against the defaults 49 / 52)
mkdir -p gpurun_out
python profiles/sweep_variants.py C3 8 49,53,49,53 > gpurun_out/r02al_prefetch_c3.log 2>&1; grep variant gpurun_out/r02al_prefetch_c3.log
python profiles/sweep_variants.py C5 2 52,54,52,54 > gpurun_out/r02al_prefetch_c5.log 2>&1; grep variant gpurun_out/r02al_prefetch_c5.log
