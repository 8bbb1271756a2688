import sys, torch
sys.path.insert(0, ".")
import __graft_entry__ as e
pkg = e.load_package()
rt = pkg.RayTracer(0)
rt.create_world(100000, 0.1); rt.build_octree(300)
nx, ny, ns = 3840, 2160, 64
fb = torch.empty((ny, nx, 3), device="cuda")
rt.render_device(rt.args(nx, ny, ns, True), fb.data_ptr())
bad = torch.nonzero(~torch.isfinite(fb).all(dim=2))
print("non-finite pixels (j, i):", bad.tolist(), [fb[j, i].tolist() for j, i in bad.tolist()])
