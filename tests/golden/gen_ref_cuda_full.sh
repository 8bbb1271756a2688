#!/bin/bash
# gen_ref_cuda_full.sh — full-size golden digests + the reference-CUDA timing record (test infrastructure).
# Runs the REFERENCE'S OWN kernels (oracle/_ref/ref_cuda_*) at BASELINE config 3 (3840x2160, 4/16/64 spp in one
# process) and config 4 (USE_FP16, 4 spp), samples the SM clock while they run, and writes digests of the frames
# (sha256 of the float / half framebuffer, sha256 of the 8-bit image in PPM order, a 1/16 x 1/16 subsample) to
# gpurun_out/golden_full/.  tests/golden/pack_ref_cuda_full.py packs them into tests/golden/ref_cuda/.
#   usage (repo root, on a B200):  bash tests/golden/gen_ref_cuda_full.sh [c3|c4|all]
set -uo pipefail
MODE="${1:-all}"
R=oracle/_ref
O=gpurun_out/golden_full
mkdir -p "$O"
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > "$O/gpu.csv" 2>&1

clocks_start() { # tag
  nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active,utilization.gpu \
             --format=csv,noheader -lms 1000 > "$O/clocks_$1.csv" 2>/dev/null &
  CLK_PID=$!
}
clocks_stop() { kill "$CLK_PID" 2>/dev/null; wait "$CLK_PID" 2>/dev/null; }

if [ "$MODE" = c3 ] || [ "$MODE" = all ]; then
  clocks_start c3
  timeout 900 "$R/ref_cuda_n100000_oct_spl300" 3840 2160 4,16,64 --fb "$O/C3_3840x2160x64.fb" > "$O/runs_c3.jsonl" 2>> "$O/stderr.log" \
    || echo "{\"failed\": \"c3\", \"rc\": $?}" >> "$O/runs_c3.jsonl"
  clocks_stop
  python tests/golden/digest_frame.py "$O/C3_3840x2160x64.fb" 3840 2160 f32 "$O/C3_3840x2160x64" && rm -f "$O/C3_3840x2160x64.fb"
fi
if [ "$MODE" = c4 ] || [ "$MODE" = all ]; then
  clocks_start c4
  timeout 600 "$R/ref_cuda_n100000_oct_spl300_fp16" 3840 2160 4 --fb "$O/C4_3840x2160x4.fb" > "$O/runs_c4.jsonl" 2>> "$O/stderr.log" \
    || echo "{\"failed\": \"c4\", \"rc\": $?}" >> "$O/runs_c4.jsonl"
  clocks_stop
  python tests/golden/digest_frame.py "$O/C4_3840x2160x4.fb" 3840 2160 f16 "$O/C4_3840x2160x4" && rm -f "$O/C4_3840x2160x4.fb"
fi
cat "$O"/runs_*.jsonl
