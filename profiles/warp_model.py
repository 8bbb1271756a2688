"""Host model of k_render's lane occupancy (analysis tool; see warp_model.cpp).
    python profiles/warp_model.py [n=100000] [spl=300] [nx=3840] [ny=2160] [ns=8] [warps=256] [tiles_per_warp=3] [per_trip=2] [density=4] [flat=1.5] [wide=0.75] [pad_reach=40]
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

FIELDS = ("rays paths vox_visits vox_nonempty cands positives filter_pass exact_accept grid_missed trips cands_lvl cands_front pass_cap_hi pass_cap_both outer active_lane_outer loop_trips "
          "trips_with_adv lanes_adv trips_with_test lanes_test trips_with_exact lanes_exact b_rounds b_chunks b_lanes_round b2_rounds b2_chunks "
          "b2_max_adv c2_chunks c2_slots_used c4_chunks c4_slots_used d2_chunks d2_rounds d4_chunks d4_rounds").split()


def main():
    a = [float(x) for x in sys.argv[1:]]
    vals = (a + [100000, 300, 3840, 2160, 8, 256, 3, 2, 4.0, 1.5, 0.75, 40.0][len(a):])[:12]
    n, spl, nx, ny, ns, warps, tpw, per_trip = [int(v) for v in vals[:8]]
    density, flat, wide, reach = float(vals[8]), float(vals[9]), float(vals[10]), float(vals[11])
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(ROOT, "gpurun_out", "libwarp_model.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-I/usr/local/cuda/include",
                    os.path.join(here, "warp_model.cpp"), "-o", so], check=True)
    lib = C.CDLL(so)
    O = entry.load_oracle()
    sph, _ = O.create_world(n)
    blob, _ = O.build_octree(sph, spl)
    cam = O.camera(nx, ny, O.ARITH_DEVICE).as_array()
    tiles = ((nx + 7) // 8) * ((ny + 3) // 4)
    step = max(tpw, tiles // warps)
    out = np.zeros(1024, dtype=np.float64)
    lib.wm_run.restype = C.c_int
    lib.hs_set_grid_shape(C.c_float(flat), C.c_float(wide))
    lib.hs_set_pad_reach(C.c_float(reach))
    nd = lib.wm_run(C.c_void_p(sph.ctypes.data), len(sph), C.c_void_p(cam.ctypes.data), C.c_void_p(blob.ctypes.data), spl, C.c_float(density), nx, ny, ns,
                    50, 0, step, warps, tpw, per_trip, C.c_void_p(out.ctypes.data), len(out))
    assert nd > 0, nd
    r = dict(zip(FIELDS, out))
    rays = r["rays"]
    print(f"rays {rays:.0f} paths {r['paths']:.0f} rays/path {rays / r['paths']:.3f}")
    for k in ("vox_visits", "vox_nonempty", "cands", "positives", "filter_pass", "exact_accept", "grid_missed", "trips", "cands_lvl", "cands_front", "pass_cap_hi", "pass_cap_both"):
        print(f"  per ray: {k:13s} {r[k] / rays:8.3f}")
    print(f"outer iterations {r['outer']:.0f}, lanes with a ray {r['active_lane_outer'] / r['outer']:.2f}")
    lt = r["loop_trips"]
    print(f"present loop: warp trips per outer {lt / r['outer']:.2f} (per ray {lt / rays:.3f}); lane-trips / (32 x warp trips) = {r['trips'] / (32 * lt):.3f}")
    print(f"  trips with an advance {r['trips_with_adv'] / lt:.3f} (lanes {r['lanes_adv'] / max(r['trips_with_adv'], 1):.2f}), "
          f"with a test {r['trips_with_test'] / lt:.3f} (lanes {r['lanes_test'] / max(r['trips_with_test'], 1):.2f}), "
          f"with an exact test {r['trips_with_exact'] / lt:.3f} (lanes {r['lanes_exact'] / max(r['trips_with_exact'], 1):.2f})")
    print(f"design B : rounds per outer {r['b_rounds'] / r['outer']:.2f}, chunks per outer {r['b_chunks'] / r['outer']:.2f}, "
          f"fill {r['cands'] / (32 * r['b_chunks']):.3f}, ray lanes per round {r['b_lanes_round'] / r['b_rounds']:.2f}")
    print(f"design B2: rounds per outer {r['b2_rounds'] / r['outer']:.2f}, chunks per outer {r['b2_chunks'] / r['outer']:.2f}, "
          f"fill {r['cands'] / (32 * r['b2_chunks']):.3f}, advance steps per round (max over lanes) {r['b2_max_adv'] / r['b2_rounds']:.2f}")
    print(f"k_render_coop: chunk steps per outer with 2 references per lane {r['c2_chunks'] / r['outer']:.2f} (candidate slots filled "
          f"{r['c2_slots_used'] / (64 * r['c2_chunks']):.3f}), with 4 per lane {r['c4_chunks'] / r['outer']:.2f} (filled {r['c4_slots_used'] / (128 * r['c4_chunks']):.3f})")
    print(f"deferred tails: 2 per lane: chunk steps per outer {r['d2_chunks'] / r['outer']:.2f}, rounds {r['d2_rounds'] / r['outer']:.2f}; "
          f"4 per lane: {r['d4_chunks'] / r['outer']:.2f}, rounds {r['d4_rounds'] / r['outer']:.2f}")
    base = len(FIELDS)
    hv, hc, ht = out[base:base + 65], out[base + 65:base + 65 + 257], out[base + 65 + 257:base + 65 + 257 + 129]
    def pct(h, name):
        c = np.cumsum(h) / h.sum()
        qs = [int(np.searchsorted(c, q)) for q in (0.25, 0.5, 0.75, 0.9, 0.99)]
        print(f"  {name}: quartiles/p90/p99 {qs}, zero {h[0] / h.sum():.3f}")
    pct(hv, "voxels visited per ray")
    pct(hc, "candidates per ray")
    pct(ht, "loop trips per ray")
    o2 = base + 65 + 257 + 129
    print(f"shade: hit {out[o2] / rays:.3f} sky {out[o2 + 1] / rays:.3f}; materials lambert/metal/glass {out[o2 + 2] / rays:.3f} {out[o2 + 3] / rays:.3f} {out[o2 + 4] / rays:.3f}")


if __name__ == "__main__":
    main()
