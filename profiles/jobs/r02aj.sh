# r02aj: from how many spheres flat voxels / the cooperative kernel pay
mkdir -p gpurun_out
python profiles/sweep_flat_threshold.py > gpurun_out/r02aj_flat_threshold.log 2>&1
cat gpurun_out/r02aj_flat_threshold.log | tail -12
