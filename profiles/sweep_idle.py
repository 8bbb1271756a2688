"""Sweep the idle-lane threshold of k_render_tree on one config.  python profiles/sweep_idle.py [C3] [spp] [t1,t2,...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as entry
from bench import CONFIGS

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
n, spl, octree, nx, ny, ns, desc = CONFIGS[cfg]
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ths = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [2, 4, 6, 8, 10, 12, 16, 20, 24, 32]
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 0
pkg = entry.load_package()
rt = pkg.RayTracer(0)
rt.create_world(n, 0.1)
print("build", rt.build_octree(spl))
fb = torch.empty((ny, nx, 3), dtype=torch.float32, device="cuda")
ref = None
for th in ths:
    best = None
    for k in range(3):
        st = rt.render_device(rt.args(nx, ny, ns, octree, idle_thresh=th, variant=variant), fb.data_ptr())
        best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
    same = True if ref is None else bool(torch.equal(ref, fb))
    if ref is None:
        ref = fb.clone()
    print(cfg, "spp", ns, "variant", variant, "idle_thresh", th, "kernel_ms", round(best["kernel_ms"], 3), "Mrays/s", round(best["rays"] / best["kernel_ms"] / 1e3, 1),
          "same_image", same)
rt.close()
