import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import __graft_entry__ as e
pkg = e.load_package()
rt = pkg.RayTracer(0)
rt.create_world(488, 0.1, precision=pkg.PREC_FP16)
fb, st = rt.render(240, 160, 4, use_octree=False, precision=pkg.PREC_FP16)
np.save("gpurun_out/ours_n488_brute_fp16_240x160x4.npy", fb)
for ns in (1,):
    fb1, _ = rt.render(240, 160, 1, use_octree=False, precision=pkg.PREC_FP16)
    np.save("gpurun_out/ours_n488_brute_fp16_240x160x1.npy", fb1)
