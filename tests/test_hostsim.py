"""CPU tier: the PRODUCT's per-lane device code (rt_shade.cuh / rt_trace.cuh / rt_build.cuh), compiled for the host by
tests/hostsim, against the oracle in device arithmetic.  Bit-exact on every pixel: the sub-grid traversal returns
the same closest hit as the reference's exhaustive leaf scan, and shading/RNG consumption is identical."""
import ctypes as C

import numpy as np
import pytest


def _run(hostsim, O, n, spl, octree, nx, ny, ns, density=4.0, spheres=None):
    sph = spheres if spheres is not None else O.create_world(n)[0]
    blob = O.build_octree(sph, spl)[0] if octree else None
    cam = O.camera(nx, ny, O.ARITH_DEVICE)
    p = O.make_params(nx, ny, ns, octree, spl, O.ARITH_DEVICE)
    ref, _, ctr = O.render(sph, cam, p, blob)
    fb = np.zeros((ny, nx, 3), np.float32)
    c = O.Counters()
    camarr = cam.as_array()
    hostsim.hs_render(C.c_void_p(sph.ctypes.data), len(sph), C.c_void_p(camarr.ctypes.data),
                      C.c_void_p(blob.ctypes.data) if blob is not None else None, C.byref(p), C.c_float(density),
                      C.c_void_p(fb.ctypes.data), None, C.byref(c), None)
    return fb, ref, c, ctr


@pytest.mark.parametrize("n,spl,octree,nx,ny,ns", [
    (488, 30, False, 96, 64, 3),
    (488, 30, True, 160, 96, 4),
    (8000, 30, True, 120, 80, 3),
    (20000, 30, True, 96, 64, 2),      # leaf buckets overflow: dropped spheres must stay invisible
    (100000, 300, True, 96, 54, 1),    # undefined slots (D3)
])
def test_device_code_matches_oracle(hostsim, O, n, spl, octree, nx, ny, ns):
    fb, ref, c, ctr = _run(hostsim, O, n, spl, octree, nx, ny, ns)
    assert c.rays == ctr["rays"]
    assert np.array_equal(fb.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("density", [0.5, 16.0])
def test_grid_resolution_does_not_change_the_image(hostsim, O, density):
    fb, ref, _, _ = _run(hostsim, O, 8000, 30, True, 96, 64, 2, density=density)
    assert np.array_equal(fb.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("flat,wide", [(1.0, 1.0), (2.5, 0.5)])
def test_voxel_shape_does_not_change_the_image(hostsim, O, flat, wide):
    """Flat voxels in slab-shaped scenes (choose_grid, from 40 k spheres) are a speed knob: cubes and very flat voxels give the
    frame of the default shape, which test_device_code_matches_oracle holds to the oracle at the same size."""
    hostsim.hs_set_grid_shape(C.c_float(flat), C.c_float(wide))
    try:
        fb, ref, c, ctr = _run(hostsim, O, 100000, 300, True, 96, 54, 1)
    finally:
        hostsim.hs_set_grid_shape(C.c_float(1.5), C.c_float(0.75))
    assert c.rays == ctr["rays"]
    assert np.array_equal(fb.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("n,spl,nx,ny,ns", [(488, 30, 120, 80, 3), (8000, 30, 96, 64, 2), (100000, 300, 160, 90, 2)])
def test_cooperative_candidate_rule_returns_the_same_minimum(hostsim, O, n, spl, nx, ny, ns):
    """k_render_coop offers a candidate only in the voxel that holds its root (exit cap), prunes by certain-hit bounds and evaluates
    exactly later, in any order.  A host emulation of that rule, under the two extreme schedules, returns trace_walk's minimum
    (and its tie flag) for every ray of these frames."""
    hostsim.hs_coop_check(1)
    try:
        fb, ref, c, ctr = _run(hostsim, O, n, spl, True, nx, ny, ns)
        rays, bad = C.c_ulonglong(0), C.c_ulonglong(0)
        hostsim.hs_coop_check_result(C.byref(rays), C.byref(bad))
    finally:
        hostsim.hs_coop_check(0)
    assert rays.value == c.rays and rays.value > 10000
    assert bad.value == 0
    assert np.array_equal(fb.view(np.uint32), ref.view(np.uint32))


def test_choose_grid_shape_rule(hostsim):
    """Cubic voxels, except slab-shaped boxes from 40 k spheres: 1.5x the layers across the slab, 0.75x the columns along it."""
    def grid(lo, hi, live):
        dims = (C.c_int * 3)()
        hostsim.hs_choose_grid.restype = C.c_uint
        v = hostsim.hs_choose_grid((C.c_float * 3)(*lo), (C.c_float * 3)(*hi), C.c_uint(live), C.c_float(4.0), dims)
        assert v == dims[0] * dims[1] * dims[2]
        return tuple(dims)
    slab_lo, slab_hi = (-11.0, 0.0, -11.0), (11.0, 0.2, 11.0)
    small = grid(slab_lo, slab_hi, 8000)                 # below the threshold: cubes
    edge = [(slab_hi[k] - slab_lo[k]) / small[k] for k in range(3)]
    assert max(edge) / min(edge) < 1.6
    cubic_edge = (22.0 * 0.2 * 22.0 / 400000.0) ** (1.0 / 3.0)
    big = grid(slab_lo, slab_hi, 100000)
    assert big[1] == int(np.ceil(np.ceil(0.2 / cubic_edge) * 1.5))
    assert big[0] == big[2] == int(np.ceil(np.ceil(22.0 / cubic_edge) * 0.75))
    cube = grid((-5.0, -5.0, -5.0), (5.0, 5.0, 5.0), 100000)      # not a slab: cubes at any size
    assert cube[0] == cube[1] == cube[2]
    thin_z = grid((-11.0, -11.0, 0.0), (11.0, 11.0, 0.2), 100000)   # the thin axis is found, not assumed to be y
    assert thin_z == (big[0], big[2], big[1])


def test_irregular_scene(hostsim, O):
    """Radii from 0.02 to 1.5, spheres poking out of the root box and fully outside it (dropped by the reference)."""
    rng = np.random.default_rng(11)
    n = 600
    sph = np.zeros(n, dtype=O.SPHERE_DTYPE)
    sph[0] = (0, -1000, -1, 1000, 0, 0.5, 0.5, 0.5, 0)
    for i in range(1, n):
        r = float(rng.choice([0.02, 0.1, 0.2, 0.5, 1.5], p=[0.3, 0.4, 0.2, 0.08, 0.02]))
        c = (rng.uniform(-12, 12), rng.uniform(0, 2.2), rng.uniform(-12, 12))
        mat = int(rng.integers(0, 3))
        sph[i] = (c[0], c[1], c[2], r, mat, rng.random(), rng.random(), rng.random(), 1.5 if mat == 2 else rng.random() * 0.5)
    sph[17]["mat"] = -1                      # an undefined slot in the middle
    fb, ref, c, ctr = _run(hostsim, O, n, 30, True, 120, 80, 3, spheres=sph)
    assert c.rays == ctr["rays"]
    assert np.array_equal(fb.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("n,spl,expect_fast", [(488, 30, True), (8000, 30, True), (100000, 300, True), (20000, 30, False)])
def test_visible_fast_never_accepts_what_the_reference_rule_rejects(hostsim, O, n, spl, expect_fast):
    """rt_trace.cuh visible_fast is a SUFFICIENT condition for the reference's visibility rule (a cell that stores the sphere
    passes the line test): on 200 k adversarial rays per scene — origins on cell planes, zero direction components, far
    origins, hits next to cell faces, grazing hits — it must never say yes where sphere_visible says no, and it should
    settle most hits.  With overflowed buckets (20 000 spheres at SPL 30) it must defer everything to the exact rule."""
    sph, _ = O.create_world(n)
    blob, _ = O.build_octree(sph, spl)
    out = np.zeros(4, dtype=np.uint64)
    hostsim.hs_visible_soundness.restype = C.c_int
    hostsim.hs_visible_soundness(C.c_void_p(sph.ctypes.data), len(sph), C.c_void_p(blob.ctypes.data), spl, C.c_uint64(1984 + n),
                                 200000, C.c_void_p(out.ctypes.data))
    hits, fast, exact, unsound = (int(x) for x in out)
    assert hits > 100000 and unsound == 0, (hits, fast, exact, unsound)
    if expect_fast:
        assert fast >= 0.85 * exact, (hits, fast, exact)
    else:
        assert fast == 0
