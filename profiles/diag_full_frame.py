"""Full-size parity diagnosis against the reference's CUDA build (run on the GPU box; test infrastructure, not product).

    python profiles/diag_full_frame.py [ns=64] [variants=1,40]

1. renders BASELINE config 3 at full size with each kernel variant, prints the digests next to the reference's
   (tests/golden/ref_cuda/manifest_full.json) and compares the 1/16 subsample pixel by pixel;
2. lists the pixels on which the variants disagree with each other or with the subsample (plus the known NaN pixel),
   renders exactly those pixels with the REFERENCE'S OWN device code (oracle/_ref/ref_cuda_* --pixels: running sums after every
   sample) and reports, per pixel and variant, the first sample at which the running sums part.
Writes gpurun_out/diag_full_frame.json and .npz.
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402
from digest_frame import digest, subsample16  # noqa: E402

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 64
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 40]
nx, ny, n, spl = 3840, 2160, 100000, 300
out_dir = os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)
man = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_cuda", "manifest_full.json")))
gold = man["frames"]["C3_3840x2160x64"]
gsub = np.load(os.path.join(ROOT, "tests", "golden", "ref_cuda", gold["subsample"]))

pkg = entry.load_package()
rt = pkg.RayTracer(0)
rt.create_world(n, 0.1)
rt.build_octree(spl)
report = {"ns": ns, "reference": gold, "variants": {}}
frames = {}
fb = torch.empty((ny, nx, 3), dtype=torch.float32, device="cuda")
for v in variants:
    st = rt.render_device(rt.args(nx, ny, ns, True, variant=v), fb.data_ptr())
    f = fb.cpu().numpy()
    frames[v] = f
    d = digest(f)
    d["kernel"], d["kernel_ms"], d["rays"] = st["kernel"], st["kernel_ms"], st["rays"]
    if ns == 64:
        sub = subsample16(f)
        neq = (sub.view(np.uint32) != gsub.view(np.uint32)).any(axis=2)
        d["subsample_pixels_differing"] = int(neq.sum())
        d["subsample_pixels"] = int(neq.size)
        d["matches_reference_sha"] = d["sha256_raw"] == gold["sha256_raw"]
    report["variants"][v] = d
    print("variant", v, json.dumps(d), flush=True)

# pixels to look at
pix = {(2070, 687)}
if ns == 64:
    for v in variants:
        sub = subsample16(frames[v])
        jj, ii = np.nonzero((sub.view(np.uint32) != gsub.view(np.uint32)).any(axis=2))
        for j, i in list(zip(jj, ii))[:64]:
            pix.add((int(i) * 16, int(j) * 16))
if len(variants) > 1:
    a, b = frames[variants[0]], frames[variants[1]]
    jj, ii = np.nonzero((a.view(np.uint32) != b.view(np.uint32)).any(axis=2))
    report["variant_pixels_differing"] = int(len(jj))
    for j, i in list(zip(jj, ii))[:160]:
        pix.add((int(i), int(j)))
pix = sorted(pix)
print("pixels under the microscope:", len(pix), flush=True)
plist = os.path.join(out_dir, "diag_pixels.txt")
with open(plist, "w") as fh:
    for i, j in pix:
        fh.write(f"{i} {j}\n")

# the reference's own device code on those pixels: running sums after every sample
ref_bin = os.path.join(ROOT, "oracle", "_ref", "ref_cuda_n100000_oct_spl300")
pout = os.path.join(out_dir, "diag_ref_prefix.bin")
r = subprocess.run([ref_bin, str(nx), str(ny), str(ns), "--pixels", plist, "--pixels-out", pout], capture_output=True, text=True, timeout=600)
print(r.stdout.strip(), r.stderr.strip()[-300:], flush=True)
ref_prefix = np.fromfile(pout, dtype=np.float32).reshape(len(pix), ns, 3)

# our running sums: the linear sums of a k-sample frame are the sums after sample k (one stream per pixel)
ii = torch.tensor([p[0] for p in pix], device="cuda")
jj = torch.tensor([p[1] for p in pix], device="cuda")
ours = {}
for v in variants:
    pre = np.zeros((len(pix), ns, 3), np.float32)
    for k in range(1, ns + 1):
        rt.render_accumulate(rt.args(nx, ny, k, True, variant=v), fb.data_ptr(), want_stats=False)
        pre[:, k - 1] = fb[jj, ii].cpu().numpy()
    ours[v] = pre
rows = []
for q, (i, j) in enumerate(pix):
    row = {"pixel": [i, j], "ref_final": ref_prefix[q, -1].tolist()}
    for v in variants:
        neq = (ours[v][q].view(np.uint32) != ref_prefix[q].view(np.uint32)).any(axis=1)
        nan_eq = np.isnan(ours[v][q]).all(axis=1) & np.isnan(ref_prefix[q]).all(axis=1)
        neq &= ~nan_eq
        row[f"v{v}_first_diff_sample"] = int(np.argmax(neq)) if neq.any() else -1
        row[f"v{v}_final"] = ours[v][q, -1].tolist()
    rows.append(row)
report["pixels"] = rows
for v in variants:
    bad = [r["pixel"] for r in rows if r[f"v{v}_first_diff_sample"] >= 0]
    print(f"variant {v}: {len(bad)} of {len(rows)} examined pixels part from the reference's running sums: {bad[:40]}", flush=True)
np.savez_compressed(os.path.join(out_dir, "diag_full_frame.npz"), pixels=np.array(pix), ref_prefix=ref_prefix,
                    **{f"v{v}_prefix": ours[v] for v in variants})
json.dump(report, open(os.path.join(out_dir, "diag_full_frame.json"), "w"), indent=1)
rt.close()
