"""Old vs pooled kernel across scene sizes (octree mode, 1920x1080, 4 spp).  python profiles/scene_sizes.py [variants]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as entry

variants = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 11]
pkg = entry.load_package()
rt = pkg.RayTracer(0)
nx, ny, ns = 1920, 1080, 4
fb = torch.empty((ny, nx, 3), dtype=torch.float32, device="cuda")
for n, spl in ((488, 30), (2000, 30), (8000, 30), (30000, 100), (100000, 300), (300000, 1000), (1000000, 3000)):
    rt.create_world(n, 0.1)
    b = rt.build_octree(spl)
    ref = None
    for v in variants:
        best = None
        for k in range(2):
            st = rt.render_device(rt.args(nx, ny, ns, True, variant=v), fb.data_ptr())
            best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
        if ref is None:
            ref = fb.clone()
        print("n", n, "variant", v, "kernel_ms", round(best["kernel_ms"], 3), "Mrays/s", round(best["rays"] / best["kernel_ms"] / 1e3, 1),
              "same", bool(torch.equal(ref, fb)), "build_ms", round(b["build_ms"], 2), "refs", b["fine_refs"], flush=True)
rt.close()
