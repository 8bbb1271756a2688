"""Host-side timeline of the octree build and of the e2e pieces (upload, build, render, D2H).   python profiles/time_build.py [C3]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
from bench import CONFIGS
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
n, spl, octree, nx, ny, ns, _ = CONFIGS[cfg]
pkg = entry.load_package()
rt = pkg.RayTracer(0)
rt.create_world(n, 0.1)
rt.build_octree(spl)
sph = rt.spheres()
for k in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); rt.upload_world(sph); torch.cuda.synchronize(); t1 = time.perf_counter()
    if k == 2: os.environ["RT_BUILD_TRACE"] = "1"
    st = rt.build_octree(spl); torch.cuda.synchronize(); t2 = time.perf_counter()
    os.environ.pop("RT_BUILD_TRACE", None)
    print(f"upload_world {1e3*(t1-t0):.3f} ms, build_octree {1e3*(t2-t1):.3f} ms (device events {st['build_ms']:.3f} ms)", flush=True)
rt.close()
