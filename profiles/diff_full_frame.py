"""Pixel-by-pixel difference between this repo's full-size C3 frame and the one the reference's CUDA build writes (run on the
GPU box, ~5 minutes of it for the reference).  Writes gpurun_out/diff_full_frame.{json,npz} and the reference frame's per-row
sha256 list (gpurun_out/golden_full/C3_3840x2160x64_rows.json).  Test infrastructure, not product."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

nx, ny, ns, n, spl = 3840, 2160, 64, 100000, 300
out = os.path.join(ROOT, "gpurun_out")
os.makedirs(os.path.join(out, "golden_full"), exist_ok=True)
fbp = "/tmp/ref_c3.fb"
r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_cuda_n100000_oct_spl300"), str(nx), str(ny), str(ns), "--fb", fbp],
                   capture_output=True, text=True, timeout=900)
print(r.stdout.strip(), flush=True)
ref = np.fromfile(fbp, dtype=np.float32).reshape(ny, nx, 3)
rows = [hashlib.sha256(ref[j].tobytes()).hexdigest()[:16] for j in range(ny)]
json.dump({"sha256_raw": hashlib.sha256(ref.tobytes()).hexdigest(), "row_sha256_16": rows},
          open(os.path.join(out, "golden_full", "C3_3840x2160x64_rows.json"), "w"))
pkg = entry.load_package()
rt = pkg.RayTracer(0)
rt.create_world(n, 0.1)
rt.build_octree(spl)
fb = torch.empty((ny, nx, 3), dtype=torch.float32, device="cuda")
st = rt.render_device(rt.args(nx, ny, ns, True), fb.data_ptr())
ours = fb.cpu().numpy()
neq = (ours.view(np.uint32) != ref.view(np.uint32)).any(axis=2)
jj, ii = np.nonzero(neq)
print("kernel", st["kernel"], "pixels differing:", len(jj), flush=True)
pix = np.stack([ii, jj], 1)
np.savez_compressed(os.path.join(out, "diff_full_frame.npz"), pixels=pix, ref=ref[jj, ii], ours=ours[jj, ii])
json.dump({"differing": int(len(jj)), "pixels": pix[:200].tolist(), "ref": ref[jj, ii][:200].tolist(), "ours": ours[jj, ii][:200].tolist(),
           "sha_ours": hashlib.sha256(ours.tobytes()).hexdigest(), "sha_ref": hashlib.sha256(ref.tobytes()).hexdigest()},
          open(os.path.join(out, "diff_full_frame.json"), "w"), indent=1)
for k in range(min(len(jj), 30)):
    print(int(ii[k]), int(jj[k]), ref[jj[k], ii[k]], ours[jj[k], ii[k]])
os.remove(fbp)
