"""CPU tier: include/rt_dropin.h — the reference's world-building class surface (sphere / hitable_list / lambertian / metal /
dielectric / camera) as host classes that flatten into the C ABI's scene description.  A C++ program builds the world
through those classes; the flattened result must equal the scene it was built from, and hitable_list::hit must pick
the sphere the oracle picks."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fmt(x):
    return float(x).hex()


def test_dropin_world_roundtrip(O, tmp_path):
    exe = str(tmp_path / "world_roundtrip")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "dropin", "world_roundtrip.cpp"), "-o", exe], check=True)
    sph, _ = O.create_world(488)
    sph = sph.copy()
    sph[7]["mat"], sph[7]["param"] = 1, 1.75            # a metal whose fuzz the constructor must clamp (material.h:66)
    rr = np.random.default_rng(2)
    Q = 200
    org = np.stack([rr.uniform(-12, 13, Q), rr.uniform(0.05, 3, Q), rr.uniform(-12, 12, Q)], 1).astype(np.float32)
    dirs = rr.normal(size=(Q, 3)).astype(np.float32)
    lines = [str(len(sph))]
    for s in sph:
        lines.append(" ".join([_fmt(s["cx"]), _fmt(s["cy"]), _fmt(s["cz"]), _fmt(s["radius"]), str(int(s["mat"])), _fmt(s["ax"]),
                               _fmt(s["ay"]), _fmt(s["az"]), _fmt(s["param"])]))
    lines.append(str(Q))
    for o, d in zip(org, dirs):
        lines.append(" ".join(_fmt(v) for v in (*o, *d)))
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.splitlines()
    want = sph.copy()
    want[7]["param"] = 1.0
    for i, s in enumerate(want):
        f = out[i].split()
        got = [float.fromhex(f[k]) for k in (0, 1, 2, 3)] + [int(f[4])] + [float.fromhex(f[k]) for k in (5, 6, 7, 8)]
        exp = [float(s["cx"]), float(s["cy"]), float(s["cz"]), float(s["radius"]), int(s["mat"]), float(s["ax"]), float(s["ay"]),
               float(s["az"]), float(s["param"])]
        if s["mat"] == 2:
            exp[5:8] = [0.0, 0.0, 0.0]
        assert got == exp, (i, got, exp)
    hits = 0
    for k in range(Q):
        idx, t = out[len(sph) + k].split()
        oi, ot = O.closest_hit(want, org[k], dirs[k], None, 30, False, O.ARITH_HOST)
        assert int(idx) == oi, (k, idx, oi)
        if oi >= 0:
            assert abs(float.fromhex(t) - ot) <= 1e-5 * max(1.0, abs(ot))
            hits += 1
    assert hits > 20
    assert out[len(sph) + Q].startswith("camera ")
