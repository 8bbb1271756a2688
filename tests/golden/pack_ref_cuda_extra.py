"""Add later golden dumps (tests/golden/gen_ref_cuda.sh upseed / fp16 modes, run on a B200) to tests/golden/ref_cuda/ without
re-packing the earlier ones.   python tests/golden/pack_ref_cuda_extra.py [gpurun_out/golden_ref_cuda]"""
import hashlib
import json
import os
import sys

import numpy as np

SRC = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden_ref_cuda"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_cuda")
mp = os.path.join(DST, "manifest.json")
manifest = json.load(open(mp))
extra = manifest.setdefault("extra_frames", {})
FRAMES = {  # file -> (n, spl, use_octree, nx, ny, ns, seed_mode, fp16)
    "n488_oct_upseed_240x160x4.fb": (488, 30, 1, 240, 160, 4, 1, 0),
    "n488_oct_fp16_240x160x4.fb": (488, 30, 1, 240, 160, 4, 0, 1),
    "n488_brute_fp16_240x160x4.fb": (488, 30, 0, 240, 160, 4, 0, 1),
    "n8000_oct_fp16_240x160x4.fb": (8000, 30, 1, 240, 160, 4, 0, 1),
    "n100000_oct_fp16_192x108x2.fb": (100000, 300, 1, 192, 108, 2, 0, 1),
}
for f, (n, spl, octree, nx, ny, ns, seed_mode, fp16) in FRAMES.items():
    p = os.path.join(SRC, f)
    if not os.path.exists(p):
        continue
    raw = np.fromfile(p, dtype="<f2" if fp16 else "<f4").reshape(ny, nx, 3)
    name = f.replace(".fb", ".npz")
    np.savez_compressed(os.path.join(DST, name), fb=raw)
    extra[name] = {"n": n, "spl": spl, "use_octree": octree, "nx": nx, "ny": ny, "ns": ns, "seed_mode": seed_mode, "fp16": fp16,
                   "sha256": hashlib.sha256(raw.tobytes()).hexdigest()}
# USE_FP16 camera (22 values widened to float by the dump kernel), scene, and the leaf lists of the FP16 Octree
for f in sorted(os.listdir(SRC)):
    if f.startswith("cam_fp16_") and f.endswith(".bin"):
        manifest.setdefault("cameras_fp16", {})[f[9:-4]] = [float(x) for x in np.fromfile(os.path.join(SRC, f), dtype="<f4")]
SPHERE_DTYPE = np.dtype([("cx", "<f4"), ("cy", "<f4"), ("cz", "<f4"), ("radius", "<f4"), ("mat", "<i4"),
                         ("ax", "<f4"), ("ay", "<f4"), ("az", "<f4"), ("param", "<f4")])
q = os.path.join(SRC, "n488_fp16.spheres")
if os.path.exists(q):
    np.save(os.path.join(DST, "n488_fp16_spheres.npy"), np.fromfile(q, dtype=SPHERE_DTYPE))


def fp16_cell_lists(path, spl):
    """USE_FP16 Octree layout: nodes[585] x 48 B {level i32, AABB 6 x f16, children[8] i32}, leaves[4097] x (spl+1) i32,
    nodeCount, leafCount.  Returns {(x_low, y_low, z_low) of a level-3 node: stored sphere indices in order}."""
    raw = open(path, "rb").read()
    nodes = np.frombuffer(raw[:585 * 48], dtype=np.uint8).reshape(585, 48)
    leaves = np.frombuffer(raw[585 * 48:585 * 48 + 4097 * (spl + 1) * 4], dtype="<i4").reshape(4097, spl + 1)
    node_count = int(np.frombuffer(raw[-8:-4], dtype="<i4")[0])
    out = {}
    for k in range(node_count):
        level = int(nodes[k, :4].view("<i4")[0])
        if level != 3:
            continue
        box = nodes[k, 4:16].view("<f2").astype(np.float32)
        ch = nodes[k, 16:48].view("<i4")
        lst = []
        for c in ch:
            if c == 0:
                break
            lst.extend(int(x) for x in leaves[c, :leaves[c, spl]])
        out[(float(box[0]), float(box[1]), float(box[2]))] = lst
    return out


for f, spl in (("n488_spl30_fp16.octree", 30), ("n8000_spl30_fp16.octree", 30)):
    q = os.path.join(SRC, f)
    if os.path.exists(q):
        cl = fp16_cell_lists(q, spl)
        keys = sorted(cl)
        np.savez_compressed(os.path.join(DST, f.replace(".octree", "_cells.npz")), low=np.array(keys, dtype=np.float32),
                            start=np.cumsum([0] + [len(cl[k]) for k in keys]).astype(np.int64),
                            idx=np.array([i for k in keys for i in cl[k]], dtype=np.int32))
runs = os.path.join(SRC, "runs.jsonl")
if os.path.exists(runs):
    have = {json.dumps(r, sort_keys=True) for r in manifest.get("runs", [])}
    for line in open(runs):
        try:
            r = json.loads(line)
        except Exception:
            continue
        if json.dumps(r, sort_keys=True) not in have:
            manifest.setdefault("runs", []).append(r)
json.dump(manifest, open(mp, "w"), indent=1)
print("extra frames:", sorted(extra))
