// rt_build.cuh — per-item logic of the parallel octree build (shared by the build kernels in rt_octree.cu).
//
// Reference semantics being reproduced (acceleration_structure.h:82-217):
//   * intersects(): box grown by the radius contains the centre; x uses (low, high], y and z use [low, high];
//   * child boxes by halving in float; child index = (x_high<<2)|(y_high<<1)|z_high   -> the path of a level-3
//     cell from the root is its 9-bit MORTON code, 3 bits per level;
//   * a sphere reaches a level-3 cell iff it passes the test at every box on the path; the tests are monotone in
//     the box, so that is the product of three per-axis index ranges at level 3;
//   * nodes and leaf buckets are numbered in creation order of the serial insertion: sorted by
//     (first sphere that touches them, pre-order position inside that sphere's insertion).
#pragma once
#include "rt_math.cuh"
#include "rt_types.h"

namespace rt {

// potential-node ids: level 0: 0; level 1: 1 + c1; level 2: 9 + c1*8 + c2; level 3: 73 + morton
RT_HD int level_base(int level) { return ((1 << (3 * level)) - 1) / 7; }   // 0, 1, 9, 73

// 3 bits per level, x bit highest: morton = sum over levels L=1..3 of child(L) << (3*(3-L))
RT_HD int morton_of(int ix, int iy, int iz) {
    int m = 0;
#pragma unroll
    for (int b = 2; b >= 0; b--) m = (m << 3) | (((ix >> b) & 1) << 2) | (((iy >> b) & 1) << 1) | ((iz >> b) & 1);
    return m;
}
RT_HD void morton_to_xyz(int m, int &ix, int &iy, int &iz) {
    ix = iy = iz = 0;
#pragma unroll
    for (int l = 0; l < 3; l++) {
        const int c = (m >> (3 * (2 - l))) & 7;
        ix = (ix << 1) | (c >> 2);
        iy = (iy << 1) | ((c >> 1) & 1);
        iz = (iz << 1) | (c & 1);
    }
}

// pre-order rank key of a potential node: digits (child+1) in base 9, parents (shorter paths) first
RT_HD int preorder_key(int level, int path /* morton prefix, 3*level bits */) {
    int key = 0;
    for (int l = 0; l < 3; l++) {
        int digit = 0;
        if (l < level) digit = ((path >> (3 * (level - 1 - l))) & 7) + 1;
        key = key * 9 + digit;
    }
    return key;
}

struct AxisRange {
    int lo, hi;   // inclusive; empty when lo > hi
};

// acceleration_structure.h:82-93 along one axis over the 8 level-3 slabs [P[i], P[i+1]]
RT_HD AxisRange axis_range(const float *P, float c, float r, bool open_low /* x axis: (low, high] */) {
    AxisRange a;
    a.lo = 8; a.hi = -1;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float low = sub_(P[i], r), high = add_(P[i + 1], r);
        const bool in = (open_low ? (c > low) : (c >= low)) && (c <= high);
        if (in) { a.lo = imin(a.lo, i); a.hi = imax(a.hi, i); }
    }
    return a;
}

// ---- fine sub-grid -------------------------------------------------------------------------------------------
// How far outside a sphere the float test of sphere.h:17-23 can still report a hit: the rounding error of
// c = dot(oc,oc) - r*r is about 3 ulp(|oc|^2), and the miss distance that produces is that error over 2r.
// kSceneReach bounds |oc| (camera at |(13,2,3)|, spheres within |x|,|z| <= 11.1).
constexpr float kSceneReach = 40.0f;
RT_HD float sphere_pad(float r, float reach = kSceneReach) {
    const float rr = fmaxf(r, 1e-3f);
    return fmaxf(2e-4f, 4e-7f * reach * reach / rr);
}

// Does the surface of sphere (c, r), thickened by pad, cross the box [lo, hi]?
RT_HD bool shell_hits_box(const float4 s, const float pad, const float *lo, const float *hi) {
    float dmin2 = 0.f, dmax2 = 0.f;
    const float c[3] = {s.x, s.y, s.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float a = lo[k] - c[k], b = hi[k] - c[k];
        const float nearest = a > 0.f ? a : (b < 0.f ? b : 0.f);
        const float farthest = fmaxf(fabsf(a), fabsf(b));
        dmin2 += nearest * nearest;
        dmax2 += farthest * farthest;
    }
    const float ro = s.w + pad, ri = fmaxf(s.w - pad, 0.f);
    return dmin2 <= ro * ro && dmax2 >= ri * ri;
}


// Grid resolution: about `density` voxels per gridded sphere.  Voxels are cubes, except for slab-shaped scenes from
// kFlatVoxelMinSpheres spheres up (the domain of the cooperative kernel, kCoopMinSpheres): the reference's scene is a carpet of
// equal spheres resting on y = 0, and there voxels `flat` times thinner across the slab and 1/`wide` times wider along it cut
// the candidates per ray by a fifth at an unchanged number of voxel steps (profiles/warp_model.py: C3 19.4 -> 16.9 candidates,
// 10.1 -> 8.8 chunk steps per warp trace; measured, profiles/r02/r02ah_shape_*.log: C3 8 spp 17.04 -> 15.79 ms, C5 2 spp
// 47.8 -> 39.5 ms with 6 % fewer references; the pixel-per-lane kernel of the small scenes does not gain).  The grid is an
// internal structure: its shape cannot change a frame (the sweep checks the hash).
// Fills org/hi/vs/inv_vs/n*; returns the voxel count (0 when there is nothing to grid).  Runs on the host.
constexpr float kGridFlat = 1.5f, kGridWide = 0.75f;
constexpr uint32_t kFlatVoxelMinSpheres = 40000;
inline uint32_t choose_grid(const float *lo, const float *hi, uint32_t live, float density, GridView &g, float flat = kGridFlat,
                            float wide = kGridWide, uint32_t flat_min = kFlatVoxelMinSpheres) {
    g.nx = g.ny = g.nz = 0;
    if (live == 0) return 0;
    float sz[3];
    for (int k = 0; k < 3; k++) {
        g.org[k] = lo[k];
        g.hi[k] = hi[k];
        sz[k] = fmaxf(hi[k] - lo[k], 1e-4f);
    }
    int thin = 0;                               // the axis across the slab, if the box is one
    for (int k = 1; k < 3; k++)
        if (sz[k] < sz[thin]) thin = k;
    bool slab = live >= flat_min;
    for (int k = 0; k < 3; k++)
        if (k != thin && sz[thin] * 4.0f > sz[k]) slab = false;
    const float vol = sz[0] * sz[1] * sz[2];
    float target = density * (float)live;
    target = fminf(fmaxf(target, 1.f), 16777216.f);
    const float edge = cbrtf(vol / target);
    int d[3];
    for (int k = 0; k < 3; k++) {
        float c = ceilf(sz[k] / edge);
        if (slab) c = ceilf(c * (k == thin ? flat : wide));
        c = fminf(fmaxf(c, 1.f), 2047.f);
        d[k] = (int)c;
        g.vs[k] = sz[k] / (float)d[k];
        g.inv_vs[k] = (float)d[k] / sz[k];
    }
    g.nx = d[0]; g.ny = d[1]; g.nz = d[2];
    return (uint32_t)d[0] * (uint32_t)d[1] * (uint32_t)d[2];
}

// voxel index range (inclusive, clamped) a padded sphere can touch along each axis
RT_HD void voxel_range(const GridView &g, const float4 s, float pad, int *v0, int *v1) {
    const int n[3] = {g.nx, g.ny, g.nz};
    const float c[3] = {s.x, s.y, s.z};
    const float r = s.w + pad;
    for (int k = 0; k < 3; k++) {
        int a = (int)floorf((c[k] - r - g.org[k]) * g.inv_vs[k]) - 1;
        int b = (int)floorf((c[k] + r - g.org[k]) * g.inv_vs[k]) + 1;
        v0[k] = a < 0 ? 0 : a;
        v1[k] = b > n[k] - 1 ? n[k] - 1 : b;
    }
}
RT_HD void voxel_box(const GridView &g, int x, int y, int z, float *lo, float *hi) {
    lo[0] = g.org[0] + (float)x * g.vs[0]; hi[0] = g.org[0] + (float)(x + 1) * g.vs[0];
    lo[1] = g.org[1] + (float)y * g.vs[1]; hi[1] = g.org[1] + (float)(y + 1) * g.vs[1];
    lo[2] = g.org[2] + (float)z * g.vs[2]; hi[2] = g.org[2] + (float)(z + 1) * g.vs[2];
}

}  // namespace rt
