import sys, torch
sys.path.insert(0, ".")
import __graft_entry__ as e
pkg = e.load_package()
rt = pkg.RayTracer(0)
for n, spl, nx, ny, ns in [(488, 30, 1200, 800, 10), (8000, 30, 3840, 2160, 8), (50000, 300, 3840, 2160, 8)]:
    rt.create_world(n, 0.1); rt.build_octree(spl); rt.set_camera(nx, ny)
    fb = torch.empty((ny, nx, 3), device="cuda"); ref = None
    for v in (22, 0, 22, 0):
        st = min((rt.render_device(rt.args(nx, ny, ns, True, variant=v), fb.data_ptr()) for _ in range(3)), key=lambda s: s["kernel_ms"])
        if ref is None: ref = fb.clone()
        print(f"n={n} {nx}x{ny}x{ns} variant {v}: {st['kernel_ms']:.3f} ms {st['rays']/st['kernel_ms']/1e3:.0f} Mrays/s same={bool(torch.equal(ref, fb))}", flush=True)
