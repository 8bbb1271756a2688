import sys, torch
sys.path.insert(0, ".")
import __graft_entry__ as e
pkg = e.load_package()
rt = pkg.RayTracer(0)
n, spl = int(sys.argv[1]), int(sys.argv[2])
rt.create_world(n, 0.1); rt.build_octree(spl)
for nx, ny, ns in [(7680, 4320, 1), (7680, 4320, 4), (3840, 2160, 16), (1920, 1080, 64), (960, 540, 256), (3840, 2160, 64)]:
    rt.set_camera(nx, ny)
    fb = torch.empty((ny, nx, 3), device="cuda")
    st = rt.render_device(rt.args(nx, ny, ns, True), fb.data_ptr())
    print(f"n={n} {nx}x{ny}x{ns}: {st['kernel_ms']:.1f} ms {st['rays']/st['kernel_ms']/1e3:.0f} Mrays/s", flush=True)
