#!/bin/bash
# gen_ref_cuda.sh — golden-vector generator (test infrastructure).
# Runs the REFERENCE'S OWN CUDA kernels (oracle/_ref/ref_cuda_*, built by `make -C oracle ref_cuda` from
# /root/reference, recompiled for sm_100) on the GPU box and dumps float framebuffers, sphere lists, cameras
# and Octree blobs into gpurun_out/golden_ref_cuda/.  tests/golden/pack_ref_cuda.py then packs them into the
# fixtures committed under tests/golden/ref_cuda/.
#   usage (from the repo root, on a B200):  bash tests/golden/gen_ref_cuda.sh [small|upseed|fp16|timing|all]
set -uo pipefail
MODE="${1:-all}"
R=oracle/_ref
O=gpurun_out/golden_ref_cuda
mkdir -p "$O"
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > "$O/gpu.csv" 2>&1
nproc > "$O/nproc.txt"; lscpu | head -20 > "$O/lscpu.txt" 2>&1

run() { # variant tag nx ny ns extra...
  local v="$1" tag="$2" nx="$3" ny="$4" ns="$5"; shift 5
  echo "== $v $tag ${nx}x${ny}x${ns}" >&2
  timeout 600 "$R/ref_cuda_$v" "$nx" "$ny" "$ns" "$@" >> "$O/runs.jsonl" 2>> "$O/stderr.log" || echo "{\"failed\": \"$v $tag\", \"rc\": $?}" >> "$O/runs.jsonl"
}

if [ "$MODE" = small ] || [ "$MODE" = all ]; then
  run n488_brute_spl30 small 240 160 4 --fb "$O/n488_brute_240x160x4.fb" --spheres "$O/n488.spheres" --camera "$O/cam_240x160.bin"
  run n488_oct_spl30   small 240 160 4 --fb "$O/n488_oct_240x160x4.fb" --octree "$O/n488_spl30.octree"
  run n8000_oct_spl30  small 240 160 4 --fb "$O/n8000_oct_240x160x4.fb" --spheres "$O/n8000.spheres" --octree "$O/n8000_spl30.octree"
  run n100000_oct_spl300 small 192 108 2 --fb "$O/n100000_oct_192x108x2.fb" --spheres "$O/n100000.spheres" --octree "$O/n100000_spl300.octree" --camera "$O/cam_192x108.bin"
  run n1000000_oct_spl3000 small 96 54 1 --fb "$O/n1000000_oct_96x54x1.fb" --camera "$O/cam_96x54.bin"
  run n488_brute_spl30 cam 1200 800 1 --camera "$O/cam_1200x800.bin"
  run n488_brute_spl30 cam 3840 2160 1 --camera "$O/cam_3840x2160.bin"
  run n488_brute_spl30 cam 7680 4320 1 --camera "$O/cam_7680x4320.bin"
  # big dumps are only checksummed (they are reproducible from the oracle)
  sha256sum "$O"/*.spheres "$O"/*.octree > "$O/sha256.txt"
  rm -f "$O/n100000.spheres" "$O/n100000_spl300.octree" "$O/n8000.spheres"
fi
if [ "$MODE" = upseed ] || [ "$MODE" = all ]; then
  # the upstream per-pixel seeding curand_init(1984, pixel_index, 0) (main.cu:90 swapped in for :93 by patch_ref.sh)
  run n488_oct_spl30_upseed upseed 240 160 4 --fb "$O/n488_oct_upseed_240x160x4.fb"
  run n488_oct_spl30_upseed upseed_timing 1200 800 10 --reps 3
fi
if [ "$MODE" = fp16 ] || [ "$MODE" = all ]; then
  # the reference built with -DUSE_FP16 (precision_types.h:8): frames are 3 halves per pixel
  run n488_brute_spl30_fp16 fp16 240 160 4 --fb "$O/n488_brute_fp16_240x160x4.fb" --spheres "$O/n488_fp16.spheres" --camera "$O/cam_fp16_240x160.bin"
  run n488_oct_spl30_fp16   fp16 240 160 4 --fb "$O/n488_oct_fp16_240x160x4.fb" --octree "$O/n488_spl30_fp16.octree"
  run n8000_oct_spl30_fp16  fp16 240 160 4 --fb "$O/n8000_oct_fp16_240x160x4.fb" --octree "$O/n8000_spl30_fp16.octree"
  run n100000_oct_spl300_fp16 fp16 192 108 2 --fb "$O/n100000_oct_fp16_192x108x2.fb"
  run n488_brute_spl30_fp16 C1_fp16 1200 800 10 --reps 3 --fb "$O/C1_fp16_1200x800x10.fb"
  run n488_oct_spl30_fp16   C2_fp16 1200 800 10 --reps 3
  run n100000_oct_spl300_fp16 C4_4spp 3840 2160 4
  sha256sum "$O"/*fp16*.octree > "$O/sha256_fp16.txt"
fi
if [ "$MODE" = timing ] || [ "$MODE" = all ]; then
  run n488_brute_spl30 C1 1200 800 10 --reps 3 --fb "$O/C1_1200x800x10.fb"
  run n488_oct_spl30   C2 1200 800 10 --reps 3 --fb "$O/C2_1200x800x10.fb"
  run n8000_oct_spl30  stock 1200 800 10 --reps 3
  run n100000_oct_spl300 C3_4spp 3840 2160 4
  run n100000_oct_spl300 C3_64spp 3840 2160 64
  run n1000000_oct_spl3000 C5_1spp 7680 4320 1
fi
cat "$O/runs.jsonl"
