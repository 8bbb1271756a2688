"""Pack the dumps written by tests/golden/gen_ref_cuda.sh (run on a B200) into the fixtures committed under
tests/golden/ref_cuda/.  The dumps come from the REFERENCE'S OWN CUDA kernels (oracle/_ref/ref_cuda_*).

    python tests/golden/pack_ref_cuda.py [gpurun_out/golden_ref_cuda]
"""
import hashlib
import json
import os
import sys

import numpy as np

SRC = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden_ref_cuda"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_cuda")
os.makedirs(DST, exist_ok=True)

SPHERE_DTYPE = np.dtype([("cx", "<f4"), ("cy", "<f4"), ("cz", "<f4"), ("radius", "<f4"), ("mat", "<i4"),
                         ("ax", "<f4"), ("ay", "<f4"), ("az", "<f4"), ("param", "<f4")])

manifest = {"source": "oracle/_ref/ref_cuda_* on NVIDIA B200 (reference kernels, nvcc 12.9 -O3 sm_100)", "frames": {},
            "cameras": {}, "sha256": {}, "runs": []}

FRAMES = {  # file -> (n, spl, use_octree, nx, ny, ns)
    "n488_brute_240x160x4.fb": (488, 30, 0, 240, 160, 4),
    "n488_oct_240x160x4.fb": (488, 30, 1, 240, 160, 4),
    "n8000_oct_240x160x4.fb": (8000, 30, 1, 240, 160, 4),
    "n100000_oct_192x108x2.fb": (100000, 300, 1, 192, 108, 2),
}
for f, (n, spl, octree, nx, ny, ns) in FRAMES.items():
    p = os.path.join(SRC, f)
    if not os.path.exists(p):
        continue
    fb = np.fromfile(p, dtype="<f4").reshape(ny, nx, 3)
    name = f.replace(".fb", ".npz")
    np.savez_compressed(os.path.join(DST, name), fb=fb)
    manifest["frames"][name] = {"n": n, "spl": spl, "use_octree": octree, "nx": nx, "ny": ny, "ns": ns,
                                "sha256_f32": hashlib.sha256(fb.tobytes()).hexdigest()}

# full-size frames of BASELINE configs 1 and 2: quantised exactly as the reference's PPM writer does (main.cu:327)
for f, tag in (("C1_1200x800x10.fb", "C1"), ("C2_1200x800x10.fb", "C2")):
    p = os.path.join(SRC, f)
    if not os.path.exists(p):
        continue
    fb = np.fromfile(p, dtype="<f4").reshape(800, 1200, 3)
    q = (255.99 * fb.astype(np.float64)).astype(np.int64)[::-1].astype(np.uint8)
    manifest["frames"][f"{tag}_1200x800x10"] = {"sha256_f32": hashlib.sha256(fb.tobytes()).hexdigest(),
                                               "sha256_u8_ppm_order": hashlib.sha256(q.tobytes()).hexdigest()}
    if tag == "C2":
        np.savez_compressed(os.path.join(DST, "C2_1200x800x10_u8.npz"), rgb=q)

for f in sorted(os.listdir(SRC)):
    if f.startswith("cam_") and f.endswith(".bin"):
        manifest["cameras"][f[4:-4]] = [float(x) for x in np.fromfile(os.path.join(SRC, f), dtype="<f4")]
        manifest["cameras"][f[4:-4] + "_hex"] = np.fromfile(os.path.join(SRC, f), dtype="<u4").tolist()

for f, key in (("n488.spheres", "n488.spheres"), ("n488_spl30.octree", "n488_spl30.octree"),
               ("n8000_spl30.octree", "n8000_spl30.octree")):
    p = os.path.join(SRC, f)
    if os.path.exists(p):
        manifest["sha256"][key] = hashlib.sha256(open(p, "rb").read()).hexdigest()
sha = os.path.join(SRC, "sha256.txt")
if os.path.exists(sha):
    for line in open(sha):
        h, name = line.split()
        manifest["sha256"][os.path.basename(name)] = h
p = os.path.join(SRC, "n488.spheres")
if os.path.exists(p):
    np.save(os.path.join(DST, "n488_spheres.npy"), np.fromfile(p, dtype=SPHERE_DTYPE))
runs = os.path.join(SRC, "runs.jsonl")
if os.path.exists(runs):
    for line in open(runs):
        try:
            manifest["runs"].append(json.loads(line))
        except Exception:
            pass
for extra in ("gpu.csv", "nproc.txt"):
    q = os.path.join(SRC, extra)
    if os.path.exists(q):
        manifest[extra] = open(q).read().strip()
json.dump(manifest, open(os.path.join(DST, "manifest.json"), "w"), indent=1)
print("packed into", DST, os.listdir(DST))
