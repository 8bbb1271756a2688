#!/bin/bash
# r02ag: voxel shape of the traversal grid (anisotropic voxels), C3 8 spp and C5 2 spp
mkdir -p gpurun_out
python profiles/sweep_grid_shape.py C3 8 "1:1,1.5:0.75,2:0.75,1.5:0.85,2:0.85,2.5:0.75,2:0.65,1.75:0.75,2.25:0.85" > gpurun_out/r02ag_shape_c3.log 2>&1
python profiles/sweep_grid_shape.py C5 2 "1:1,1.5:0.75,2:0.75,1.5:0.85,1.25:0.75,1.5:0.65" > gpurun_out/r02ag_shape_c5.log 2>&1
cat gpurun_out/r02ag_shape_c3.log gpurun_out/r02ag_shape_c5.log
