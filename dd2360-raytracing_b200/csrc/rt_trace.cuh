// rt_trace.cuh — closest-hit queries: flat list (hitable_list.h:16-31) and octree (acceleration_structure.h:226-342).
//
// Result contract: the closest hit is the minimum over a CANDIDATE SET, with strict '<', so it does not
// depend on the order candidates are tested in, nor on testing candidates that cannot win.
//   flat list : candidate set = all spheres.
//   octree    : candidate set = ground sphere (index 0) + the stored lists of every level-3 cell whose reference
//               AABB passes the reference's infinite-LINE slab test (intersect_ray_aabb, :226-244, evaluated here
//               with the same float operations).  A parent box contains its children and float subtraction and
//               division are monotone, so a passing cell implies passing ancestors.
// What this file is free to do, and does: visit cells front to back, skip cells and voxels the RAY cannot reach
// before the current closest hit, and look only at the spheres whose surface crosses the voxels the ray walks
// (a per-cell uniform sub-grid, 3D-DDA).  All pruning is conservative (margins below), so the minimum is unchanged.
#pragma once
#include "rt_shade.cuh"

namespace rt {

#if defined(__CUDA_ARCH__)
#define RT_LDG(p) __ldg(p)
#else
#define RT_LDG(p) (*(p))
#endif

// A float-evaluated root can sit this far (relative, along the ray) from where exact arithmetic would put it:
// near-tangent hits lose half the mantissa in sqrt(b*b - a*c).  Pruning keeps this much slack.
constexpr float kTSlackRel = 4e-3f;
constexpr float kTSlackAbs = 1e-3f;

struct Hit {
    float t;
    int idx;
};

// Work counters of the instrumented build (-DRT_COUNTERS -> librt_b200_counters.so); compiled out otherwise.
struct TraceCounters {
    uint32_t sphere_tests, node_tests, voxel_steps;
};
#ifdef RT_COUNTERS
#define RT_COUNT(field) (tc.field++)
#else
#define RT_COUNT(field) ((void)0)
#endif

// Ties: two DIFFERENT spheres with bit-identical t keep whichever is tested first, here as in the reference
// (strict '<').  Voxel reference lists are sorted at build time, so the outcome is deterministic run to run;
// it could differ from the reference's pick only for such exact ties, which the generated scenes do not produce.

// ---- flat list: every sphere, SoA float4, warp-uniform address (broadcast) ------------------------------------
template <typename GeomPtr>
RT_HD Hit trace_list(const GeomPtr geom, const int *tag, const int n, const vec3f o, const vec3f d, TraceCounters &tc) {
    Hit h;
    h.t = kTMax; h.idx = -1;
    const float a = dot3(d, d);
    for (int i = 0; i < n; i++) {
        const float4 s = geom[i];
        float t;
        RT_COUNT(sphere_tests);
        // slots create_world never wrote carry radius 0 and tag NONE: skipped
        if (sphere_test(s, o, d, a, h.t, t) && RT_LDG(tag + i) >= 0) { h.t = t; h.idx = i; }
    }
    return h;
}

// acceleration_structure.h:226-244 — same operations, same comparisons (NaN compares false as there)
RT_HD bool ref_line_test(const vec3f o, const vec3f d, const float xl, const float yl, const float zl, const float xh,
                         const float yh, const float zh) {
    float tmin = div_(sub_(xl, o.x), d.x), tmax = div_(sub_(xh, o.x), d.x);
    if (tmin > tmax) { const float t = tmin; tmin = tmax; tmax = t; }
    float tymin = div_(sub_(yl, o.y), d.y), tymax = div_(sub_(yh, o.y), d.y);
    if (tymin > tymax) { const float t = tymin; tymin = tymax; tymax = t; }
    if ((tmin > tymax) || (tymin > tmax)) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = div_(sub_(zl, o.z), d.z), tzmax = div_(sub_(zh, o.z), d.z);
    if (tzmin > tzmax) { const float t = tzmin; tzmin = tzmax; tzmax = t; }
    if ((tmin > tzmax) || (tzmin > tmax)) return false;
    return true;
}

struct RayPre {
    vec3f o, d, inv;    // inv = 1/d (IEEE; +-inf for zero components)
    float a;            // dot(d,d)
};

// Conservative ray/box interval: returns false when the ray cannot be inside [lo,hi] for any t in
// (kTMin, t_far].  fminf/fmaxf drop NaNs (0*inf), which only ever widens the interval.
RT_HD bool ray_box(const RayPre &r, const float *lo, const float *hi, const float t_far, float &t_enter, float &t_exit) {
    const float tx0 = (lo[0] - r.o.x) * r.inv.x, tx1 = (hi[0] - r.o.x) * r.inv.x;
    const float ty0 = (lo[1] - r.o.y) * r.inv.y, ty1 = (hi[1] - r.o.y) * r.inv.y;
    const float tz0 = (lo[2] - r.o.z) * r.inv.z, tz1 = (hi[2] - r.o.z) * r.inv.z;
    float t0 = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), 0.0f));
    float t1 = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), t_far));
    t_enter = t0;
    t_exit = t1;
    // slack: the boxes are padded at build time; this covers the rounding of the products above
    return t0 <= t1 * (1.0f + 1e-5f) + 1e-6f;
}

// Walk one cell's sub-grid with a 3D-DDA and test the spheres registered in every voxel the ray crosses.
RT_HD void trace_cell(const SceneView &sc, const TreeView &tv, const CellGrid &g, const RayPre &r, float t0, float t1,
                      Hit &h, TraceCounters &tc) {
    const int nx = (int)(g.dims & 1023u), ny = (int)((g.dims >> 10) & 1023u), nz = (int)(g.dims >> 20);
    // entry point, nudged inside; clamp handles rounding at the faces
    const float te = fmaxf(t0, 0.0f);
    int ix = (int)floorf((r.o.x + r.d.x * te - g.org[0]) * g.inv_vs[0]);
    int iy = (int)floorf((r.o.y + r.d.y * te - g.org[1]) * g.inv_vs[1]);
    int iz = (int)floorf((r.o.z + r.d.z * te - g.org[2]) * g.inv_vs[2]);
    ix = imin(imax(ix, 0), nx - 1);
    iy = imin(imax(iy, 0), ny - 1);
    iz = imin(imax(iz, 0), nz - 1);
    const int sx = r.d.x >= 0.0f ? 1 : -1, sy = r.d.y >= 0.0f ? 1 : -1, sz = r.d.z >= 0.0f ? 1 : -1;
    // parameter at which the ray leaves the current voxel along each axis
    float tmx = (g.org[0] + (float)(ix + (sx > 0)) * g.vs[0] - r.o.x) * r.inv.x;
    float tmy = (g.org[1] + (float)(iy + (sy > 0)) * g.vs[1] - r.o.y) * r.inv.y;
    float tmz = (g.org[2] + (float)(iz + (sz > 0)) * g.vs[2] - r.o.z) * r.inv.z;
    const float dtx = fabsf(g.vs[0] * r.inv.x), dty = fabsf(g.vs[1] * r.inv.y), dtz = fabsf(g.vs[2] * r.inv.z);
    // a zero direction component never advances that axis
    if (!(fabsf(r.d.x) > 0.0f)) tmx = kTMax;
    if (!(fabsf(r.d.y) > 0.0f)) tmy = kTMax;
    if (!(fabsf(r.d.z) > 0.0f)) tmz = kTMax;
    float t_in = te;
    const int max_steps = nx + ny + nz + 3;
    for (int step = 0; step < max_steps; step++) {
        // a later voxel can only hold hits at t >= t_in (minus the float slack)
        if (t_in > h.t * (1.0f + kTSlackRel) + kTSlackAbs) break;
        RT_COUNT(voxel_steps);
        const uint32_t v = g.vox_base + (uint32_t)((iz * ny + iy) * nx + ix);
        const uint32_t b = RT_LDG(tv.vox_start + v), e = RT_LDG(tv.vox_start + v + 1);
        for (uint32_t k = b; k < e; k++) {
            const int idx = (int)RT_LDG(tv.vox_refs + k);
            const float4 s = RT_LDG(sc.geom + idx);
            float t;
            RT_COUNT(sphere_tests);
            if (sphere_test(s, r.o, r.d, r.a, h.t, t)) { h.t = t; h.idx = idx; }
        }
        // step to the neighbour the ray enters next
        if (tmx <= tmy && tmx <= tmz) { t_in = tmx; ix += sx; tmx += dtx; if ((unsigned)ix >= (unsigned)nx) break; }
        else if (tmy <= tmz)          { t_in = tmy; iy += sy; tmy += dty; if ((unsigned)iy >= (unsigned)ny) break; }
        else                          { t_in = tmz; iz += sz; tmz += dtz; if ((unsigned)iz >= (unsigned)nz) break; }
        if (t_in > t1 * (1.0f + 1e-5f) + 1e-6f) break;
    }
}

// acceleration_structure.h:319-342 hitTree, re-organised (see the header comment)
// `planes` = the 3 x 9 slab plane coordinates (kept in the kernel's constant parameter space)
RT_HD Hit trace_tree(const SceneView &sc, const TreeView &tv, const float *planes, const vec3f o, const vec3f d,
                    TraceCounters &tc) {
    Hit h;
    h.t = kTMax; h.idx = -1;
    RayPre r;
    r.o = o; r.d = d;
    r.a = dot3(d, d);
    r.inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    {   // ground sphere first (:322-332)
        float t;
        RT_COUNT(sphere_tests);
        if (sphere_test(RT_LDG(sc.geom), o, d, r.a, kTMax, t)) { h.t = t; h.idx = 0; }
    }
    // front-to-back child order: flip the child bits along which the ray travels in the negative direction
    const int flip = (d.x < 0.0f ? 4 : 0) | (d.y < 0.0f ? 2 : 0) | (d.z < 0.0f ? 1 : 0);
    const TreeNode &root = tv.nodes[0];
    for (int k1 = 0; k1 < 8; k1++) {
        const int c1 = root.child[k1 ^ flip];
        if (c1 == 0) continue;   // 0 = absent, as in the reference (the root is nobody's child)
        float te, tx;
        RT_COUNT(node_tests);
        if (!ray_box(r, tv.node_ext[c1].lo, tv.node_ext[c1].hi, h.t * (1.0f + kTSlackRel) + kTSlackAbs, te, tx)) continue;
        const TreeNode &n1 = tv.nodes[c1];
        for (int k2 = 0; k2 < 8; k2++) {
            const int c2 = n1.child[k2 ^ flip];
            if (c2 == 0) continue;
            RT_COUNT(node_tests);
            if (!ray_box(r, tv.node_ext[c2].lo, tv.node_ext[c2].hi, h.t * (1.0f + kTSlackRel) + kTSlackAbs, te, tx)) continue;
            const TreeNode &n2 = tv.nodes[c2];
            for (int k3 = 0; k3 < 8; k3++) {
                const int c3 = n2.child[k3 ^ flip];
                if (c3 == 0) continue;
                const TreeNode &n3 = tv.nodes[c3];
                if (n3.first_cell == 0xffffffffu) continue;   // nothing traceable stored in this cell
                const int cell = (int)n3.first_cell;
                RT_COUNT(node_tests);
                if (!ray_box(r, tv.cell_ext[cell].lo, tv.cell_ext[cell].hi, h.t * (1.0f + kTSlackRel) + kTSlackAbs, te, tx))
                    continue;
                // the reference only looks into this cell when the infinite line crosses its AABB
                RT_COUNT(node_tests);
                if (!ref_line_test(o, d, planes[n3.ix], planes[kPlanes + n3.iy], planes[2 * kPlanes + n3.iz],
                                   planes[n3.ix + 1], planes[kPlanes + n3.iy + 1], planes[2 * kPlanes + n3.iz + 1]))
                    continue;
                const CellGrid &g = tv.cells[cell];
                // big spheres of this cell: tested directly
                const uint32_t nb = g.big & 0xffu, bb = g.big >> 8;
                for (uint32_t k = 0; k < nb; k++) {
                    const int idx = (int)RT_LDG(tv.big_refs + bb + k);
                    float t;
                    RT_COUNT(sphere_tests);
                    if (sphere_test(RT_LDG(sc.geom + idx), o, d, r.a, h.t, t)) { h.t = t; h.idx = idx; }
                }
                // small spheres: walk the sub-grid where the ray overlaps it
                if (g.dims && ray_box(r, g.org, g.hi, h.t * (1.0f + kTSlackRel) + kTSlackAbs, te, tx))
                    trace_cell(sc, tv, g, r, te, tx, h, tc);
            }
        }
    }
    return h;
}

}  // namespace rt
