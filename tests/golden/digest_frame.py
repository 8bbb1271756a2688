"""digest_frame.py — digests of one raw framebuffer dump (test infrastructure; used by gen_ref_cuda_full.sh on the GPU box
and by the GPU parity tests on the product's own frame, so both sides are reduced by the same code).

    python tests/golden/digest_frame.py frame.fb nx ny f32|f16 out_prefix

Writes out_prefix.json {sha256_raw, sha256_u8_ppm_order, nonfinite, mean} and out_prefix_sub16.npy (every 16th pixel of
every 16th row, raw dtype) — a full 4K frame (99.5 MB) does not fit the 64 MiB that travel back from the box.
"""
from __future__ import annotations

import hashlib
import json
import sys

import numpy as np


def quantise_ppm_order(fb: np.ndarray) -> np.ndarray:
    """main.cu:321-333: rows top to bottom (j = ny-1 .. 0), int(255.99 * channel) with C's truncating conversion."""
    x = 255.99 * fb[::-1].astype(np.float64)
    with np.errstate(invalid="ignore"):
        q = np.where(np.isfinite(x), np.trunc(np.clip(x, -2147483648.0, 2147483647.0)), -2147483648.0)
    return q.astype(np.int64).astype(np.int32)


def digest(fb: np.ndarray) -> dict:
    """fb: (ny, nx, 3) float32 or float16, row j = image row j of the reference's framebuffer."""
    f32 = fb.astype(np.float32)
    return {
        "sha256_raw": hashlib.sha256(np.ascontiguousarray(fb).tobytes()).hexdigest(),
        "sha256_i32_ppm_order": hashlib.sha256(np.ascontiguousarray(quantise_ppm_order(f32)).tobytes()).hexdigest(),
        "nonfinite_pixels": int((~np.isfinite(f32)).any(axis=2).sum()),
        "mean_finite": float(np.nanmean(np.where(np.isfinite(f32), f32, np.nan).astype(np.float64))),
        "dtype": str(fb.dtype), "ny": int(fb.shape[0]), "nx": int(fb.shape[1]),
    }


def subsample16(fb: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(fb[::16, ::16])


def main() -> None:
    path, nx, ny, kind, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
    dt = np.float32 if kind == "f32" else np.float16
    fb = np.fromfile(path, dtype=dt).reshape(ny, nx, 3)
    d = digest(fb)
    with open(out + ".json", "w") as f:
        json.dump(d, f, indent=1)
    np.save(out + "_sub16.npy", subsample16(fb))
    print(json.dumps(d))


if __name__ == "__main__":
    main()
