#!/bin/bash
# r02ah: voxel shape, finer sweep + density interplay + the small scene (k_render)
mkdir -p gpurun_out
python profiles/sweep_grid_shape.py C3 8 "1:1,1.5:0.65,1.75:0.65,1.6:0.7,1.75:0.7,2:0.7,1.75:0.8,1.6:0.75,1.75:0.75:3,1.75:0.75:5,1.75:0.75:6,1.5:0.75:6,1.25:0.65:8" > gpurun_out/r02ah_shape_c3.log 2>&1
python profiles/sweep_grid_shape.py C5 2 "1:1,1.6:0.75,1.75:0.75,1.75:0.65,1.5:0.55,1.6:0.65,1.5:0.7,1.5:0.75:3,1.5:0.75:5,1.5:0.75:6" > gpurun_out/r02ah_shape_c5.log 2>&1
python profiles/sweep_grid_shape.py C2 10 "1:1,1.5:0.75,1.75:0.75,2:0.75,1.6:0.75" > gpurun_out/r02ah_shape_c2.log 2>&1
cat gpurun_out/r02ah_shape_c3.log gpurun_out/r02ah_shape_c5.log gpurun_out/r02ah_shape_c2.log
