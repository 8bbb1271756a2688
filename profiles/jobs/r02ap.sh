# r02ap: A/B 9 blocks per SM (56 registers; variants 53 / 54) against the defaults (8 blocks, 64 registers; 49 / 52)
mkdir -p gpurun_out
python profiles/sweep_variants.py C3 8 49,53,49,53 > gpurun_out/r02ap_blocks9_c3.log 2>&1; grep variant gpurun_out/r02ap_blocks9_c3.log
python profiles/sweep_variants.py C5 2 52,54,52,54 > gpurun_out/r02ap_blocks9_c5.log 2>&1; grep variant gpurun_out/r02ap_blocks9_c5.log
