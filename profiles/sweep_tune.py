"""Sweep the pool kernel's TEST stickiness knobs.  python profiles/sweep_tune.py [C3] [spp] [variant]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as entry
from bench import CONFIGS

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
n, spl, octree, nx, ny, ns, desc = CONFIGS[cfg]
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 8
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 11
pkg = entry.load_package()
rt = pkg.RayTracer(0)
rt.create_world(n, 0.1)
rt.build_octree(spl)
fb = torch.empty((ny, nx, 3), dtype=torch.float32, device="cuda")
for sticky in (1, 2, 4, 8, 16):
    for smin in (4, 8, 12, 16, 20, 24):
        best = None
        for k in range(2):
            st = rt.render_device(rt.args(nx, ny, ns, octree, variant=variant, max_rounds=5000000, tune=(sticky, smin)), fb.data_ptr())
            best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
        print(cfg, "spp", ns, "variant", variant, "sticky", sticky, "min", smin, "kernel_ms", round(best["kernel_ms"], 3), "Mrays/s",
              round(best["rays"] / best["kernel_ms"] / 1e3, 1), flush=True)
        if sticky == 1:
            break
rt.close()
