#!/bin/bash
# patch_ref.sh — ORACLE support (test infrastructure, NOT product code).
# Makes a THROW-AWAY patched copy of the reference sources for oracle/Makefile.  Nothing it writes is committed:
# the copy lives under oracle/_ref/.src_<kind>/ during one compile and is deleted by the Makefile afterwards.
#
# usage: patch_ref.sh <reference dir> <out dir> host|cuda
#
# Patches (each is a one-line substitution; `grep -c` below fails the build if a pattern stops matching):
#   both : main.cu:22   NUM_SPHERES 8000        -> RTO_N            (constants are not -D overridable)
#          main.cu:24   #define USE_OCTREE      -> only #ifdef RTO_USE_OCTREE
#          acceleration_structure.h:15 SPHERES_PER_LEAF 30 -> RTO_SPL
#   host : material.h:33 RANDVEC3 and camera.h:15: the three / two curand_uniform calls inside one constructor
#          argument list are sequenced left to right, which is what nvcc does for device code; g++ evaluates them
#          right to left and would render a different image;
#          counting hooks at sphere.h:18, acceleration_structure.h:227, main.cu:48 (no effect on results).
set -euo pipefail
REF="$1"; OUT="$2"; KIND="$3"
rm -rf "$OUT"; mkdir -p "$OUT"
cp "$REF"/main.cu "$REF"/*.h "$OUT"/

must() { # file pattern
  if ! grep -qE "$2" "$1"; then echo "patch_ref.sh: pattern not found in $1: $2" >&2; exit 1; fi
}

must "$OUT/main.cu" '^#define NUM_SPHERES 8000'
sed -i -E 's/^#define NUM_SPHERES 8000.*/#define NUM_SPHERES RTO_N/' "$OUT/main.cu"
must "$OUT/main.cu" '^#define USE_OCTREE'
sed -i -E 's/^#define USE_OCTREE.*/#ifdef RTO_USE_OCTREE\n#define USE_OCTREE\n#endif/' "$OUT/main.cu"
must "$OUT/acceleration_structure.h" '^#define SPHERES_PER_LEAF 30'
sed -i -E 's/^#define SPHERES_PER_LEAF 30.*/#define SPHERES_PER_LEAF RTO_SPL/' "$OUT/acceleration_structure.h"

# optional (variant suffix _upseed): the upstream per-pixel seeding curand_init(1984, pixel_index, 0) that HEAD keeps
# as a comment on main.cu:90 replaces the HEAD form on main.cu:93
if [ "${RTO_UPSEED:-0}" = 1 ]; then
  must "$OUT/main.cu" '^[[:space:]]*// curand_init\(1984, pixel_index, 0, &rand_state\[pixel_index\]\);'
  must "$OUT/main.cu" '^[[:space:]]*curand_init\(1984 \+ pixel_index, 0, 0, &rand_state\[pixel_index\]\);'
  sed -i -E 's|^([[:space:]]*)// curand_init\(1984, pixel_index, 0, &rand_state\[pixel_index\]\);|\1curand_init(1984, pixel_index, 0, \&rand_state[pixel_index]);|' "$OUT/main.cu"
  sed -i -E 's|^([[:space:]]*)curand_init\(1984 \+ pixel_index, 0, 0, &rand_state\[pixel_index\]\);|\1// (HEAD form disabled for this variant)|' "$OUT/main.cu"
fi

if [ "$KIND" = host ]; then
  must "$OUT/material.h" '^#define RANDVEC3 vec3\(curand_uniform'
  sed -i -E 's/^#define RANDVEC3 vec3\(curand_uniform.*/static inline __host__ __device__ vec3 rto_randvec3(curandState *s_) { real_t a_ = curand_uniform(s_); real_t b_ = curand_uniform(s_); real_t c_ = curand_uniform(s_); return vec3(a_, b_, c_); }\n#define RANDVEC3 rto_randvec3(local_rand_state)/' "$OUT/material.h"
  must "$OUT/camera.h" 'p = real_t\(2\.0f\)\*vec3\(curand_uniform\(local_rand_state\),curand_uniform\(local_rand_state\),0\) - vec3\(1,1,0\);'
  sed -i -E 's/p = real_t\(2\.0f\)\*vec3\(curand_uniform\(local_rand_state\),curand_uniform\(local_rand_state\),0\) - vec3\(1,1,0\);/{ real_t a_ = curand_uniform(local_rand_state); real_t b_ = curand_uniform(local_rand_state); p = real_t(2.0f)*vec3(a_,b_,0) - vec3(1,1,0); }/' "$OUT/camera.h"
  must "$OUT/sphere.h" '^    vec3 oc = r\.origin\(\) - center;'
  sed -i -E 's/^    vec3 oc = r\.origin\(\) - center;/    RTO_COUNT_SPHERE vec3 oc = r.origin() - center;/' "$OUT/sphere.h"
  must "$OUT/acceleration_structure.h" 'float tmin = \(box\.x_low - r\.origin\(\)\.x\(\)\) / r\.direction\(\)\.x\(\);'
  sed -i -E 's/float tmin = \(box\.x_low - r\.origin\(\)\.x\(\)\) \/ r\.direction\(\)\.x\(\);/RTO_COUNT_AABB float tmin = (box.x_low - r.origin().x()) \/ r.direction().x();/' "$OUT/acceleration_structure.h"
  must "$OUT/main.cu" '^        bool hit_anything = false;'
  sed -i -E '0,/^        bool hit_anything = false;/s//        RTO_COUNT_RAY bool hit_anything = false;/' "$OUT/main.cu"
fi
