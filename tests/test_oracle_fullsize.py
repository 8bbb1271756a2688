"""CPU tier: the oracle (and the product's per-lane code, via tests/hostsim) against pixels of the FULL-SIZE frame the reference's
CUDA build wrote on a B200 (tests/golden/ref_cuda/manifest_full.json, C3_tie_pixels.json): subsample pixels spread over the frame,
and the 21 pixels whose paths contain an exact tie between two different spheres stored in different level-3 cells — the cases
that pin the reference's test order (rt_trace.cuh tie_key)."""
import ctypes as C
import json
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def c3(O):
    sph, _ = O.create_world(100000)
    blob, _ = O.build_octree(sph, 300)
    return sph, blob, O.camera(3840, 2160, O.ARITH_DEVICE)


def _oracle_pixel(O, c3, i, j):
    sph, blob, cam = c3
    fb, _, _ = O.render(sph, cam, O.make_params(3840, 2160, 64, True, 300, O.ARITH_DEVICE, window=(i, i + 1, j, j + 1)), blob)
    return fb[j, i]


def test_oracle_matches_the_reference_cuda_frame_on_subsample_pixels(O, c3, golden_dir):
    man = json.load(open(os.path.join(golden_dir, "ref_cuda", "manifest_full.json")))["frames"]["C3_3840x2160x64"]
    gsub = np.load(os.path.join(golden_dir, "ref_cuda", man["subsample"]))
    assert gsub.shape == (135, 240, 3) and man["nonfinite_pixels"] == 0
    for sj, si in [(3, 7), (40, 120), (66, 201), (90, 33), (110, 150), (131, 239), (75, 129), (20, 60)]:
        got = _oracle_pixel(O, c3, 16 * si, 16 * sj)
        assert np.array_equal(got.view(np.uint32), gsub[sj, si].view(np.uint32)), (si, sj, got, gsub[sj, si])


def test_tie_pixels_follow_the_reference_test_order(O, c3, hostsim, golden_dir):
    ties = json.load(open(os.path.join(golden_dir, "ref_cuda", "C3_tie_pixels.json")))
    want = [np.array([float.fromhex(v) for v in px], dtype=np.float32) for px in ties["ref_gamma_f32_hex"]]
    assert len(want) == 21
    for (i, j), w in zip(ties["pixels"], want):
        got = _oracle_pixel(O, c3, i, j)
        assert np.array_equal(got.view(np.uint32), w.view(np.uint32)), (i, j)
    # the product's own closest-hit code (rt_trace.cuh: tie detection in the walk, finish_hit, tie_key) on four of them
    sph, blob, cam = c3
    camarr = cam.as_array()
    for (i, j), w in list(zip(ties["pixels"], want))[::6]:
        p = O.make_params(3840, 2160, 64, True, 300, O.ARITH_DEVICE, window=(i, i + 1, j, j + 1))
        fb = np.zeros((2160, 3840, 3), np.float32)
        c = O.Counters()
        hostsim.hs_render(C.c_void_p(sph.ctypes.data), len(sph), C.c_void_p(camarr.ctypes.data), C.c_void_p(blob.ctypes.data), C.byref(p),
                          C.c_float(4.0), C.c_void_p(fb.ctypes.data), None, C.byref(c), None)
        assert np.array_equal(fb[j, i].view(np.uint32), w.view(np.uint32)), (i, j)


def test_uninitialised_refracted_pixel_is_finite(O, c3):
    """Pixel (2070, 687): total internal reflection meets curand_uniform == 1.0 at sample 36 and the reference reads its uninitialised
    `refracted` (material.h:88,109); with the value its sm_100 build holds there the pixel is finite (its bits are asserted against
    the GPU frame, and the GPU frame's sha256 against the reference's, in the -m gpu tier)."""
    px = _oracle_pixel(O, c3, 2070, 687)
    assert np.isfinite(px).all() and [float(v).hex() for v in px] == ["0x1.8e109e0000000p-3", "0x1.752e2a0000000p-2", "0x1.9c17440000000p-2"]
