"""Voxel shape of the traversal grid in slab-shaped scenes (choose_grid's `flat` and `wide`, knob RT_GRID_SHAPE="flat:wide", with
RT_GRID_DENSITY): kernel time and frame hash per shape.  The grid is an internal structure: every shape must give the same frame.
    python profiles/sweep_grid_shape.py C3|C5 [spp] [shapes "ym:xz[:density],..."]"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as entry

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
n, spl, nx, ny, ns = {"C3": (100000, 300, 3840, 2160, 8), "C5": (1000000, 3000, 7680, 4320, 2), "C2": (488, 30, 1200, 800, 10)}[cfg]
if len(sys.argv) > 2:
    ns = int(sys.argv[2])
shapes = sys.argv[3] if len(sys.argv) > 3 else "1:1,1.5:0.75,2:0.75,1.5:0.85,2:0.85,2.5:0.75,2:0.65"
pkg = entry.load_package()
fb = torch.empty((ny, nx, 3), dtype=torch.float32, device="cuda")
first, rt = None, None
for sh in shapes.split(","):
    ym, xz, dens = (sh.split(":") + ["4"])[:3]
    if rt is not None:                    # the knobs are read when the context is created
        rt.close()
    os.environ["RT_GRID_DENSITY"], os.environ["RT_GRID_SHAPE"] = dens, f"{ym}:{xz}"
    rt = pkg.RayTracer(0)
    rt.create_world(n, 0.1)
    rt.set_camera(nx, ny)
    st = rt.build_octree(spl)
    best = min(rt.render_device(rt.args(nx, ny, ns, True), fb.data_ptr())["kernel_ms"] for _ in range(3))
    torch.cuda.synchronize()
    h = hashlib.sha256(fb.cpu().numpy().tobytes()).hexdigest()[:16]
    first = first or h
    print(f"{cfg} {ns} spp  flat {ym} wide {xz} density {dens}: {best:9.3f} ms  voxels {st.get('fine_voxels')} refs {st.get('fine_refs')}  frame {h} "
          f"{'same' if h == first else 'DIFFERENT'}", flush=True)
rt.close()
