#!/usr/bin/env python
"""analysis/run_experiment.py — the reference's experiment sweep (analysis/run_experiment.sh + evaluations.ipynb) over the C ABI.

The reference builds one executable per (mode, NUM_SPHERES, radius) — RayTracing_{BASELINE,OCTREE}_{488,1000..9000}_{01,02}
(run_experiment.sh:4-23) — runs each 5 times in mode 3 under `ncu --section LaunchStats/SpeedOfLight/...`, samples
`nvidia-smi memory.used` every 200 ms, keeps output.ppm and exports the ncu CSV (run_experiment.sh:31-57); the notebook then
plots frame time against the sphere count for both modes and the share of the frame spent in the render kernel.
NUM_SPHERES / USE_OCTREE / SPHERE_RADIUS are run-time arguments here, so the same grid is ONE process:

    python analysis/run_experiment.py [--sizes 488,1000,...,9000] [--radii 0.1,0.2] [--iters 5] [--nx 1200 --ny 800 --ns 10]
                                      [--out experiments] [--ppm] [--ncu]

Per run it records what the reference's harness extracts: wall time of scene generation, octree build and the
render_init+render bracket ("took X seconds", main.cu:419-432), the kernel time from CUDA events, rays, Mrays/s and the
device memory in use; `--ppm` keeps the image (output_to_stream format), `--ncu` re-runs each configuration once under ncu
with the reference's section list.  Writes <out>/runs.csv (one row per run) and <out>/summary.md (median per
configuration, octree-vs-baseline speed-up and the kernel's share of the frame — the notebook's two figures as tables).
Needs a CUDA device: the product has no CPU path.
"""
from __future__ import annotations

import argparse
import csv
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NCU_SECTIONS = ["LaunchStats", "SpeedOfLight", "MemoryWorkloadAnalysis", "ComputeWorkloadAnalysis", "Occupancy"]   # run_experiment.sh:35-36
FIELDS = ["mode", "n", "radius", "iter", "spl", "world_s", "octree_s", "render_s", "kernel_ms", "total_s", "rays", "mrays_s",
          "mem_used_mib", "ppm"]


def spheres_per_leaf(n: int) -> int:
    """SPHERES_PER_LEAF of the reference's builds: 30 (main.cu:25) — raised like its large-scene configs so that no cell overflows."""
    return 30 if n <= 10000 else (300 if n <= 100000 else 3000)


def run_one(pkg, rt, torch, mode: str, n: int, radius: float, it: int, nx: int, ny: int, ns: int, out: str, keep_ppm: bool) -> dict:
    octree = mode == "OCTREE"
    spl = spheres_per_leaf(n)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rt.create_world(n, radius)                               # rand_init + create_world (main.cu:388-401)
    rt.synchronize()
    t1 = time.perf_counter()
    if octree:
        rt.build_octree(spl)                                 # D2H + buildOctree + H2D (main.cu:405-415)
        rt.synchronize()
    t2 = time.perf_counter()
    fb, st = rt.render(nx, ny, ns, use_octree=octree)        # render_init + render + readback (main.cu:419-432)
    t3 = time.perf_counter()
    free, total = torch.cuda.mem_get_info()
    row = {"mode": mode, "n": n, "radius": radius, "iter": it, "spl": spl if octree else "", "world_s": t1 - t0, "octree_s": t2 - t1,
           "render_s": t3 - t2, "kernel_ms": st["kernel_ms"], "total_s": t3 - t0, "rays": st["rays"],
           "mrays_s": st["rays"] / st["kernel_ms"] / 1e3, "mem_used_mib": (total - free) / 2**20, "ppm": ""}
    if keep_ppm:
        d = os.path.join(out, f"RayTracing_{mode}_{n}_{int(round(radius * 10)):02d}")
        os.makedirs(d, exist_ok=True)
        row["ppm"] = os.path.join(d, f"{os.path.basename(d)}_{it}.ppm")
        with open(row["ppm"], "wb") as f:
            f.write(pkg.format_ppm(fb))
    return row


def summarise(rows: list[dict], nx: int, ny: int, ns: int) -> str:
    med = {}
    for key in sorted({(r["radius"], r["n"], r["mode"]) for r in rows}):
        sel = [r for r in rows if (r["radius"], r["n"], r["mode"]) == key]
        med[key] = {f: statistics.median(r[f] for r in sel) for f in ("world_s", "octree_s", "render_s", "kernel_ms", "total_s", "mrays_s", "mem_used_mib")}
    lines = [f"# Experiment summary — {nx}x{ny}, {ns} spp, median of {max(r['iter'] for r in rows)} runs", ""]
    for radius in sorted({k[0] for k in med}):
        lines += [f"## SPHERE_RADIUS = {radius}", "",
                  "| spheres | BASELINE kernel ms | OCTREE kernel ms | octree speed-up | BASELINE Mrays/s | OCTREE Mrays/s | octree build ms | "
                  "kernel share of frame (BASELINE / OCTREE) | memory MiB |", "|---|---|---|---|---|---|---|---|---|"]
        for n in sorted({k[1] for k in med if k[0] == radius}):
            b, o = med.get((radius, n, "BASELINE")), med.get((radius, n, "OCTREE"))
            f = lambda m, k, s=1.0, p=3: "—" if m is None else f"{m[k] * s:.{p}f}"
            share = lambda m: "—" if m is None else f"{100 * m['kernel_ms'] / 1e3 / m['total_s']:.0f} %"
            sp = "—" if not (b and o) else f"{b['kernel_ms'] / o['kernel_ms']:.2f}x"
            lines.append(f"| {n} | {f(b, 'kernel_ms')} | {f(o, 'kernel_ms')} | {sp} | {f(b, 'mrays_s', 1, 0)} | {f(o, 'mrays_s', 1, 0)} | "
                         f"{f(o, 'octree_s', 1e3, 2)} | {share(b)} / {share(o)} | {f(o or b, 'mem_used_mib', 1, 0)} |")
        lines.append("")
    return "\n".join(lines)


def main() -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--sizes", default="488,1000,2000,3000,4000,5000,6000,7000,8000,9000")     # run_experiment.sh:4-23
    ap.add_argument("--radii", default="0.1,0.2")
    ap.add_argument("--modes", default="BASELINE,OCTREE")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--nx", type=int, default=1200)
    ap.add_argument("--ny", type=int, default=800)
    ap.add_argument("--ns", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "experiments"))
    ap.add_argument("--ppm", action="store_true", help="keep every frame as <out>/RayTracing_<MODE>_<N>_<RR>/..._<iter>.ppm")
    ap.add_argument("--ncu", action="store_true", help="additionally profile each configuration once under ncu (reference section list)")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args()

    import torch

    import __graft_entry__ as entry
    if not torch.cuda.is_available():
        raise SystemExit("run_experiment.py needs a CUDA device (there is no CPU path)")
    pkg = entry.load_package()
    rt = pkg.RayTracer(args.device)
    os.makedirs(args.out, exist_ok=True)
    sizes = [int(x) for x in args.sizes.split(",") if x]
    radii = [float(x) for x in args.radii.split(",") if x]
    modes = [m for m in args.modes.split(",") if m]
    rt.create_world(488, 0.1)                         # warm the context (module load, allocations) outside the measurements
    rt.render(64, 48, 1, use_octree=False)
    rows = []
    for radius in radii:
        for n in sizes:
            for mode in modes:
                for it in range(1, args.iters + 1):
                    rows.append(run_one(pkg, rt, torch, mode, n, radius, it, args.nx, args.ny, args.ns, args.out, args.ppm))
                r = rows[-1]
                print(f"RayTracing_{mode}_{n}_{int(round(radius * 10)):02d}: kernel {r['kernel_ms']:.3f} ms, {r['mrays_s']:.0f} Mrays/s, "
                      f"frame {r['total_s'] * 1e3:.1f} ms", flush=True)
                if args.ncu:
                    d = os.path.join(args.out, f"RayTracing_{mode}_{n}_{int(round(radius * 10)):02d}")
                    os.makedirs(d, exist_ok=True)
                    cmd = ["ncu", *sum((["--section", s] for s in NCU_SECTIONS), []), "--clock-control", "none", "--export", os.path.join(d, "ncu"), "--force-overwrite",
                           sys.executable, os.path.abspath(__file__), "--sizes", str(n), "--radii", str(radius), "--modes", mode, "--iters", "1",
                           "--nx", str(args.nx), "--ny", str(args.ny), "--ns", str(args.ns), "--out", os.path.join(d, "under_ncu")]
                    subprocess.run(cmd, check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                    with open(os.path.join(d, "ncu.csv"), "w") as f:
                        subprocess.run(["ncu", "--csv", "--import", os.path.join(d, "ncu.ncu-rep")], check=False, stdout=f, stderr=subprocess.DEVNULL)
    rt.close()
    with open(os.path.join(args.out, "runs.csv"), "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=FIELDS)
        w.writeheader()
        w.writerows(rows)
    md = summarise(rows, args.nx, args.ny, args.ns)
    with open(os.path.join(args.out, "summary.md"), "w") as f:
        f.write(md + "\n")
    print(md)
    return 0


if __name__ == "__main__":
    sys.exit(main())
