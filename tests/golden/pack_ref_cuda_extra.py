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
