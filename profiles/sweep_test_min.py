"""Pool scheduler: hold TEST back until `test_min` contexts wait in it (fuller TEST rounds vs more rounds elsewhere).
    python profiles/sweep_test_min.py [C3] [spp]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as entry
from bench import CONFIGS

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
n, spl, octree, nx, ny, ns, desc = CONFIGS[cfg]
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pkg = entry.load_package()
rt = pkg.RayTracer(0)
rt.create_world(n, 0.1)
rt.build_octree(spl)
fb = torch.empty((ny, nx, 3), dtype=torch.float32, device="cuda")
ref = None
tmins = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [-1, 12, 16, 20, 24, 28, 32, 40]      # -1: off (0 selects the default, 24)
for tmin in tmins:
    best = min((rt.render_device(rt.args(nx, ny, ns, octree, variant=11, test_min=tmin), fb.data_ptr()) for _ in range(2)), key=lambda s: s["kernel_ms"])
    if ref is None:
        ref = fb.clone()
    print(cfg, "spp", ns, "test_min", tmin, "kernel_ms", round(best["kernel_ms"], 3), "Mrays/s", round(best["rays"] / best["kernel_ms"] / 1e3, 1),
          "identical", bool(torch.equal(ref, fb)), flush=True)
rt.close()
