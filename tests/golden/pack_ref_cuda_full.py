"""pack_ref_cuda_full.py — packs what gen_ref_cuda_full.sh produced on the GPU box (gpurun_out/golden_full/) into the committed
fixtures: digests of the reference CUDA build's FULL-SIZE frames of BASELINE configs 3 and 4 (sha256 of the raw frame and of the
integer PPM values, count of non-finite pixels, mean), a 1/16 x 1/16 subsample of each frame, and the timing record with clocks.
    python tests/golden/pack_ref_cuda_full.py
"""
import csv
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = os.path.join(ROOT, "gpurun_out", "golden_full")
DST = os.path.join(ROOT, "tests", "golden", "ref_cuda")


def clocks(path):
    sm, power, reasons = [], [], set()
    for r in csv.reader(open(path)):
        if len(r) < 6 or "100" not in r[5]:
            continue                                    # samples under load only
        sm.append(float(r[1].split()[0])); power.append(float(r[3].split()[0])); reasons.add(r[4].strip())
    sm.sort()
    return {"samples_under_load": len(sm), "sm_mhz_median": sm[len(sm) // 2] if sm else None, "sm_mhz_min": sm[0] if sm else None,
            "power_w_max": max(power) if power else None, "throttle_reason_masks": sorted(reasons)}


out = {"source": "oracle/_ref/ref_cuda_n100000_oct_spl300[_fp16] (the reference's own kernels, nvcc 12.9 -O3 sm_100) on NVIDIA B200, "
                 "tests/golden/gen_ref_cuda_full.sh", "frames": {}, "timing": {}}
for tag, name in (("C3", "C3_3840x2160x64"), ("C4", "C4_3840x2160x4")):
    d = json.load(open(os.path.join(SRC, name + ".json")))
    shutil.copy(os.path.join(SRC, name + "_sub16.npy"), os.path.join(DST, name + "_sub16.npy"))
    d["subsample"] = name + "_sub16.npy"
    out["frames"][name] = d
    runs = [json.loads(l) for l in open(os.path.join(SRC, f"runs_{tag.lower()}.jsonl")) if l.strip()]
    out["timing"][tag] = {"runs": [{k: r[k] for k in ("ns", "render_ms", "render_init_ms", "create_world_ms", "octree_build_host_ms") if k in r} for r in runs],
                          "clocks": clocks(os.path.join(SRC, f"clocks_{tag.lower()}.csv"))}
json.dump(out, open(os.path.join(DST, "manifest_full.json"), "w"), indent=1)
print(json.dumps(out["timing"], indent=1))
