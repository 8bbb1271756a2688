"""Host emulation of k_render_coop's candidate rule (tests/hostsim coop_walk_host: exit cap, certain-hit bounds, deferred exact tests, two
extreme schedules) against trace_walk on every ray of larger frames than the CPU tests take.  Needs tests/hostsim/libhostsim.so (built by the
CPU tests).  Last run: 4.8 M rays over 20 k / 100 k / 1 M spheres, 0 mismatches.
    python profiles/stress_coop_rule.py"""
import sys, ctypes as C, numpy as np, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as entry
O = entry.load_oracle()
hs = C.CDLL(os.path.join(ROOT, 'tests', 'hostsim', 'libhostsim.so'))
def run(n, spl, nx, ny, ns):
    sph = O.create_world(n)[0]
    blob = O.build_octree(sph, spl)[0]
    cam = O.camera(nx, ny, O.ARITH_DEVICE)
    p = O.make_params(nx, ny, ns, True, spl, O.ARITH_DEVICE)
    fb = np.zeros((ny, nx, 3), np.float32); c = O.Counters(); camarr = cam.as_array()
    hs.hs_coop_check(1)
    t0 = time.time()
    hs.hs_render(C.c_void_p(sph.ctypes.data), len(sph), C.c_void_p(camarr.ctypes.data), C.c_void_p(blob.ctypes.data), C.byref(p), C.c_float(4.0),
                 C.c_void_p(fb.ctypes.data), None, C.byref(c), None)
    rays, bad = C.c_ulonglong(0), C.c_ulonglong(0)
    hs.hs_coop_check_result(C.byref(rays), C.byref(bad)); hs.hs_coop_check(0)
    print(f"n={n} spl={spl} {nx}x{ny}x{ns}: rays {rays.value} mismatches {bad.value} ({time.time()-t0:.0f} s)", flush=True)
run(100000, 300, 960, 540, 4)
run(1000000, 3000, 240, 135, 2)
run(20000, 30, 320, 180, 4)
