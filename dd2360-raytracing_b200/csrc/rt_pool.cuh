// rt_pool.cuh — octree-mode render kernel: warp-resident path pools with state-bucketed scheduling.
// (included by rt_render.cu after c_camera and item_to_pixel)
//
// Why: one pixel per lane running "trace, then shade" bounce by bounce leaves a warp at 8/32 active lanes
// (profiles/README.md r01c): every ray needs a different number of candidate tests, only some lanes are in the voxel
// advance, the IEEE sqrt/div of a real hit, a given material, a rejection-sampling retry ... and SIMT serialises all
// of those.  Here a warp owns a POOL of NC path contexts (NC = 64 / 96 / 128, i.e. 2 - 4 per lane) kept in shared
// memory, SoA (field-major) so that lane-to-context access is bank-conflict free in the common case.  Every path is
// in exactly one STATE; the warp repeatedly
//   1. counts its contexts per state (packed byte counters, three warp-wide REDUX adds),
//   2. picks the state with the most waiting contexts,
//   3. hands up to 32 contexts of THAT state to its lanes (ballot / popc ranking through a 32-word staging row),
//   4. runs the one body of that state, converged, and writes each context's next state.
// A context is a pixel's sample chain (all samples of a pixel draw from ONE XORWOW stream — SURVEY D8 — so a pixel is
// inherently serial; the parallelism is across pixels).  States:
//   TEST    up to 4 candidate spheres of the current list: discriminant + conservative root pre-filter
//   CAND    a candidate passed the filter: the reference's exact IEEE root evaluation (sphere.h:23-44)
//   ENTER   prolog list (ground + big spheres) done: clip the ray to the grid, set up the 3D-DDA
//   STEP    voxel list done: DDA step to the next voxel
//   END     walk finished with a hit: the reference's cell-visibility rule on the winner (rt_trace.cuh)
//   DIFF / DIEL  lambertian+metal / dielectric scatter (material.h:52-116), start of the next walk
//   SAMPLE  sky colour or black, accumulate, next camera sample or pixel write-out (main.cu:101-116)
//   CLAIM   pop the next pixel from the global queue (ballot/popc-compacted atomic), seed its stream (main.cu:93)
// Closest-hit semantics are those of rt_trace.cuh (candidate set, strict '<' minimum, order-invariant), the arithmetic
// is rt_shade.cuh's; only the ORDER in which independent work is executed differs, so frames stay bit-identical.
#pragma once

namespace pool {

enum : int { S_TEST = 0, S_CAND, S_ENTER, S_STEP, S_END, S_DIFF, S_DIEL, S_SAMPLE, S_DONE };
enum : int {
    F_OX = 0, F_OY, F_OZ, F_DX, F_DY, F_DZ,          // current ray
    F_HT, F_HIDX,                                    // closest candidate so far (F_HIDX: -1 none, -2 "sample ended black")
    F_K, F_E,                                        // unread references [k, e) of the current list; bit 31 of e: prolog list
    F_TMX, F_TMY, F_TMZ, F_DTX, F_DTY, F_DTZ, F_TEXIT, F_IX, F_IY, F_IZ,   // 3D-DDA
    F_AX, F_AY, F_AZ, F_CX, F_CY, F_CZ,              // path attenuation, pixel colour sum
    F_R0, F_R1, F_R2, F_R3, F_R4, F_R5,              // XORWOW state
    F_PIX, F_SD,                                     // pixel index; sample number (24 bits) | depth << 24
    F_CM,                                            // CAND: bits 0-3 = chunk positions still to evaluate exactly, bits 4-6 = chunk length
    F_CW0, F_CW1,                                    // the state, stored as its contribution to the warp's packed per-state byte
                                                     // counters: states 0-3 -> 1 << 8*s in CW0, states 4-7 -> 1 << 8*(s-4) in CW1
    NF
};
constexpr uint32_t kPrologBit = 0x80000000u;
constexpr uint32_t kTieBit = 0x40000000u;    // F_HIDX of a hit (>= 0): two different spheres share the closest root; body_end ranks them
constexpr int kChunk = 4;
constexpr int kSticky = 4;        // TEST: chunks a lane may run back to back on one context before the warp re-schedules
constexpr int kStickyMin = 20;    // ... as long as this many lanes are still testing

template <int NC>
struct Ctx {
    uint32_t *w;
    __device__ __forceinline__ float f(int field) const { return __uint_as_float(w[field * NC]); }
    __device__ __forceinline__ uint32_t u(int field) const { return w[field * NC]; }
    __device__ __forceinline__ int i(int field) const { return (int)w[field * NC]; }
    __device__ __forceinline__ void sf(int field, float v) const { w[field * NC] = __float_as_uint(v); }
    __device__ __forceinline__ void su(int field, uint32_t v) const { w[field * NC] = v; }
    __device__ __forceinline__ vec3f v3(int field) const { return mk(f(field), f(field + 1), f(field + 2)); }
    __device__ __forceinline__ void sv3(int field, const vec3f v) const { sf(field, v.x); sf(field + 1, v.y); sf(field + 2, v.z); }
    __device__ __forceinline__ void set_state(int s) const {
        w[F_CW0 * NC] = s < 4 ? 1u << (8 * s) : 0u;
        w[F_CW1 * NC] = (s >= 4 && s < 8) ? 1u << (8 * (s - 4)) : 0u;
    }
    __device__ __forceinline__ void set_state_walk(int s) const { w[F_CW0 * NC] = 1u << (8 * s); }   // among TEST..STEP (CW1 is already 0)
    __device__ __forceinline__ void load_rng(xorwow &r) const {
        r.d = u(F_R0); r.v0 = u(F_R1); r.v1 = u(F_R2); r.v2 = u(F_R3); r.v3 = u(F_R4); r.v4 = u(F_R5);
    }
    __device__ __forceinline__ void store_rng(const xorwow &r) const {
        su(F_R0, r.d); su(F_R1, r.v0); su(F_R2, r.v1); su(F_R3, r.v2); su(F_R4, r.v3); su(F_R5, r.v4);
    }
};

__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// (the conservative root pre-filter `maybe_hit` lives in rt_trace.cuh: the pixel-per-lane kernel uses it too)

template <int NC>
__device__ __forceinline__ void begin_walk(const Ctx<NC> c, const RenderLaunch &p) {
    c.sf(F_HT, kTMax);
    c.su(F_HIDX, (uint32_t)-1);
    if (p.tree.nprolog == 1) {          // no big spheres: the ground test rides along with ENTER
        c.set_state(S_ENTER);
    } else {
        c.su(F_K, 0u);
        c.su(F_E, (uint32_t)p.tree.nprolog | kPrologBit);
        c.set_state(S_TEST);
    }
}

// ---- TEST: chunks of kChunk candidates of the current list -------------------------------------------------------
// Entered by all 32 lanes (`have` = this lane was handed a context).  A lane keeps its context in registers and runs up
// to kSticky chunks back to back while at least kStickyMin lanes are still testing: that amortises the scheduling round
// and the shared-memory traffic over the most frequent body.  Within a chunk the four tests are independent (ILP).
template <int NC>
__device__ __forceinline__ void body_test(const Ctx<NC> c, const bool have, const RenderLaunch &p, TraceCounters &tc) {
    vec3f o = mk(0, 0, 0), d = mk(0, 0, 1);
    float ht = 0.f;
    uint32_t k = 0, e = 1;
    bool pro = false;
    if (have) {
        o = c.v3(F_OX); d = c.v3(F_DX);
        ht = c.f(F_HT);
        k = c.u(F_K);
        const uint32_t eraw = c.u(F_E);
        e = eraw & ~kPrologBit;
        pro = (eraw & kPrologBit) != 0;
    }
    const float a = dot3(d, d), ia = rcp_fast(a);
    const float4 *list = pro ? p.tree.prolog_geom : p.tree.grid.ref_geom;       // geometry in list order: one dependent load
    const uint32_t last = e - 1;
    bool act = have;
    uint32_t cm = 0;
#pragma unroll 1
    for (int it = 0; it < p.tune_sticky; it++) {
        if (act) {
            const uint32_t n = min(e - k, (uint32_t)kChunk);
            float4 s[kChunk];
#pragma unroll
            for (int j = 0; j < kChunk; j++) s[j] = __ldg(list + min(k + j, last));
            uint32_t m = 0;
#pragma unroll
            for (int j = 0; j < kChunk; j++) m |= (uint32_t)maybe_hit(s[j], o, d, a, ia, ht) << j;
            m &= (1u << n) - 1u;
#ifdef RT_COUNTERS
            tc.sphere_tests += n - __popc(m);
#endif
            if (m) { cm = m | n << 4; act = false; }          // exact evaluation needed: CAND takes over at this chunk
            else { k += n; act = k < e; }
        }
        if (__popc(__ballot_sync(0xffffffffu, act)) < p.tune_sticky_min) break;
    }
    if (have) {
        c.su(F_K, k);
        if (cm) c.su(F_CM, cm);
        c.set_state_walk(cm ? S_CAND : (k < e ? S_TEST : (pro ? S_ENTER : S_STEP)));
    }
}

// ---- CAND: exact root evaluation of one flagged candidate of the chunk at k (sphere.h:17-46 via sphere_test) -----------
template <int NC>
__device__ __forceinline__ void body_cand(const Ctx<NC> c, const RenderLaunch &p, TraceCounters &tc) {
    const vec3f o = c.v3(F_OX), d = c.v3(F_DX);
    const float a = dot3(d, d);
    const float ht = c.f(F_HT);
    uint32_t k = c.u(F_K), cm = c.u(F_CM);
    const uint32_t eraw = c.u(F_E), e = eraw & ~kPrologBit;
    const bool pro = (eraw & kPrologBit) != 0;
    const int j = __ffs(cm & 15u) - 1;
    const int idx = (int)__ldg((pro ? p.tree.prolog : p.tree.grid.refs) + k + j);
    const float4 s = __ldg(p.scene.geom + idx);
    float t;
    RT_COUNT(sphere_tests);
    if (sphere_test(s, o, d, a, tie_bound(ht), t)) {
        const uint32_t hraw = c.u(F_HIDX);
        if (t < ht) { c.sf(F_HT, t); c.su(F_HIDX, (uint32_t)idx); }
        else if ((int)hraw >= 0 && (hraw & ~kTieBit) != (uint32_t)idx) c.su(F_HIDX, hraw | kTieBit);   // same root, another sphere (rt_trace.cuh "Ties")
    }
    cm &= cm - 1u;                                   // clear the lowest flagged position
    if (cm & 15u) { c.su(F_CM, cm); return; }        // more flagged candidates in this chunk: stay in CAND
    k += cm >> 4;
    c.su(F_K, k);
    c.set_state_walk(k < e ? S_TEST : (pro ? S_ENTER : S_STEP));
}

template <int NC>
__device__ __forceinline__ void load_voxel(const Ctx<NC> c, const GridView &g, int ix, int iy, int iz, TraceCounters &tc) {
    RT_COUNT(voxel_steps);
    const uint2 v = __ldg(g.vox + ((size_t)(iz * g.ny + iy) * g.nx + ix));
    c.su(F_K, v.x);
    c.su(F_E, v.x + v.y);
    c.set_state_walk(v.y ? S_TEST : S_STEP);
}

template <int NC>
__device__ __forceinline__ void end_walk(const Ctx<NC> c) {
    const bool hit = c.i(F_HIDX) >= 0;
    c.su(F_CW0, 0u);
    c.su(F_CW1, hit ? 1u << (8 * (S_END - 4)) : 1u << (8 * (S_SAMPLE - 4)));
}

// ---- ENTER: clip the ray to the grid, set up the DDA (rt_trace.cuh trace_walk, same expressions) --------------------
template <int NC>
__device__ __forceinline__ void body_enter(const Ctx<NC> c, const RenderLaunch &p, TraceCounters &tc) {
    const GridView &g = p.tree.grid;
    const vec3f o = c.v3(F_OX), d = c.v3(F_DX);
    float ht = c.f(F_HT);
    if (p.tree.nprolog == 1) {          // the ground sphere, unconditionally and first (hitTree :322-332)
        float t;
        RT_COUNT(sphere_tests);
        if (sphere_test(__ldg(p.scene.geom), o, d, dot3(d, d), kTMax, t)) { ht = t; c.sf(F_HT, t); c.su(F_HIDX, 0u); }
    }
    if (g.nx == 0) { end_walk(c); return; }
    RayPre r;
    r.o = o; r.d = d; r.a = 0.f;
    r.inv = mk(rcp_trav(d.x), rcp_trav(d.y), rcp_trav(d.z));
    float te, tx;
    if (!ray_box(r, g.org, g.hi, ht * (1.0f + kTSlackRel) + kTSlackAbs, te, tx)) { end_walk(c); return; }
    int ix = (int)floorf((o.x + d.x * te - g.org[0]) * g.inv_vs[0]);
    int iy = (int)floorf((o.y + d.y * te - g.org[1]) * g.inv_vs[1]);
    int iz = (int)floorf((o.z + d.z * te - g.org[2]) * g.inv_vs[2]);
    ix = imin(imax(ix, 0), g.nx - 1);
    iy = imin(imax(iy, 0), g.ny - 1);
    iz = imin(imax(iz, 0), g.nz - 1);
    // ray parameter at which the ray leaves the current voxel along each axis; an axis the ray does not move along
    // never advances
    c.sf(F_TMX, fabsf(d.x) > 0.0f ? (g.org[0] + (float)(ix + (d.x >= 0.0f)) * g.vs[0] - o.x) * r.inv.x : kTMax);
    c.sf(F_TMY, fabsf(d.y) > 0.0f ? (g.org[1] + (float)(iy + (d.y >= 0.0f)) * g.vs[1] - o.y) * r.inv.y : kTMax);
    c.sf(F_TMZ, fabsf(d.z) > 0.0f ? (g.org[2] + (float)(iz + (d.z >= 0.0f)) * g.vs[2] - o.z) * r.inv.z : kTMax);
    c.sf(F_DTX, fabsf(g.vs[0] * r.inv.x));
    c.sf(F_DTY, fabsf(g.vs[1] * r.inv.y));
    c.sf(F_DTZ, fabsf(g.vs[2] * r.inv.z));
    c.sf(F_TEXIT, tx);
    c.su(F_IX, (uint32_t)ix); c.su(F_IY, (uint32_t)iy); c.su(F_IZ, (uint32_t)iz);
    load_voxel(c, g, ix, iy, iz, tc);
}

// ---- STEP: into the neighbour voxel the ray enters next ------------------------------------------------------------
template <int NC>
__device__ __forceinline__ void body_step(const Ctx<NC> c, const RenderLaunch &p, TraceCounters &tc) {
    const GridView &g = p.tree.grid;
    const float tmx = c.f(F_TMX), tmy = c.f(F_TMY), tmz = c.f(F_TMZ);
    int ix = c.i(F_IX), iy = c.i(F_IY), iz = c.i(F_IZ);
    float t_in;
    bool in;
    if (tmx <= tmy && tmx <= tmz) { t_in = tmx; ix += c.f(F_DX) >= 0.0f ? 1 : -1; c.sf(F_TMX, tmx + c.f(F_DTX)); c.su(F_IX, (uint32_t)ix); in = (unsigned)ix < (unsigned)g.nx; }
    else if (tmy <= tmz)          { t_in = tmy; iy += c.f(F_DY) >= 0.0f ? 1 : -1; c.sf(F_TMY, tmy + c.f(F_DTY)); c.su(F_IY, (uint32_t)iy); in = (unsigned)iy < (unsigned)g.ny; }
    else                          { t_in = tmz; iz += c.f(F_DZ) >= 0.0f ? 1 : -1; c.sf(F_TMZ, tmz + c.f(F_DTZ)); c.su(F_IZ, (uint32_t)iz); in = (unsigned)iz < (unsigned)g.nz; }
    // a later voxel can only hold hits at t >= t_in (minus the float slack); also stop at the grid exit.  The voxel
    // index moves monotonically along each axis, so the walk leaves the grid after at most nx + ny + nz steps.
    if (!in || t_in > c.f(F_HT) * (1.0f + kTSlackRel) + kTSlackAbs || t_in > c.f(F_TEXIT) * (1.0f + 1e-5f) + 1e-6f) { end_walk(c); return; }
    load_voxel(c, g, ix, iy, iz, tc);
}

// ---- END: the reference's visibility rule, once, on the winner (rt_trace.cuh trace_tree) ---------------------------
template <int NC>
__device__ __forceinline__ void body_end(const Ctx<NC> c, const RenderLaunch &p, const float *planes, TraceCounters &tc) {
    int hidx = c.i(F_HIDX);
    if (hidx >= 0) {
        const bool tie = ((uint32_t)hidx & kTieBit) != 0u;
        Hit h;
        h.t = c.f(F_HT); h.idx = (int)((uint32_t)hidx & ~kTieBit);
        // ties, then the reference's visibility rule on the winner (rt_trace.cuh finish_hit); the rare failures redo the walk per lane
        if (tie || (h.idx > 0 && p.tree.check_visibility)) {
            const Hit w = finish_hit(p.scene, p.tree, planes, c.v3(F_OX), c.v3(F_DX), h, tie, tc);
            if (w.idx != hidx || w.t != h.t) { c.sf(F_HT, w.t); c.su(F_HIDX, (uint32_t)w.idx); }
            h = w;
        }
        hidx = h.idx;
    }
    c.set_state(hidx < 0 ? S_SAMPLE : (__ldg(p.scene.tag + hidx) == 2 /* RT_MAT_DIELECTRIC */ ? S_DIEL : S_DIFF));
}

// ---- DIFF / DIEL: material scatter (rt_shade.cuh scatter), then the next walk or the end of the sample ---------------
template <int NC>
__device__ __forceinline__ void body_shade(const Ctx<NC> c, const RenderLaunch &p, uint32_t &nrays) {
    const vec3f o = c.v3(F_OX), d = c.v3(F_DX);
    const int hidx = c.i(F_HIDX);
    const float4 g = __ldg(p.scene.geom + hidx);
    const float4 m = __ldg(p.scene.matl + hidx);
    const int tag = __ldg(p.scene.tag + hidx);
    xorwow rng;
    c.load_rng(rng);
    vec3f hp, hn, a, dn;
    hit_point(g, o, d, c.f(F_HT), hp, hn);
    const bool scattered = scatter(tag, m, d, hp, hn, a, dn, rng);
    c.store_rng(rng);
    const uint32_t sd = c.u(F_SD);
    const int depth = (int)(sd >> 24) + 1;
    if (scattered && depth < p.max_depth) {          // next iteration of color()'s loop (main.cu:47-66)
        const vec3f att = c.v3(F_AX);
        c.sv3(F_AX, mk(mul_(att.x, a.x), mul_(att.y, a.y), mul_(att.z, a.z)));
        c.sv3(F_OX, hp);
        c.sv3(F_DX, dn);
        c.su(F_SD, sd + (1u << 24));
        nrays++;
        begin_walk(c, p);
    } else {                                         // absorbed (main.cu:64) or depth exhausted (main.cu:74): black
        c.su(F_HIDX, (uint32_t)-2);
        c.set_state(S_SAMPLE);
    }
}

// new camera sample of pixel `pix` (main.cu:104-106), start of its first walk
template <int NC>
__device__ __forceinline__ void gen_sample(const Ctx<NC> c, const RenderLaunch &p, const int pix, xorwow &rng, uint32_t &nrays, uint32_t &npaths) {
    const int pj = pix / p.nx, pi = pix - pj * p.nx;
    const float u = div_(add_((float)pi, xorwow_uniform(rng)), (float)p.nx);
    const float v = div_(add_((float)pj, xorwow_uniform(rng)), (float)p.ny);
    vec3f o, d;
    camera_ray(p.cam, u, v, rng, o, d);
    c.sv3(F_OX, o);
    c.sv3(F_DX, d);
    c.sv3(F_AX, mk(1.0f, 1.0f, 1.0f));
    npaths++;
    nrays++;
    begin_walk(c, p);
}

// ---- SAMPLE: a sample ended (sky or black): accumulate, next sample — or pixel write-out and the claim of the next pixel ----
// (main.cu:68-71,107-115).  Entered by all 32 lanes (`have` = this lane was handed a context) because the queue pop is
// compacted per warp with ballot/popc: one atomic for all lanes whose pixel just finished.  F_PIX < 0 marks a context
// that has no pixel yet (start of the kernel, or a tile position outside the image).
// Pixels are claimed in whole tiles PER WARP (`stock`: the unclaimed rest of this warp's last batch; `batch` = 32 or 64 queue
// items = one or two 8x4 tiles): a context that finishes its pixel continues with the next pixel of its warp's own tile,
// so the contexts of a warp stay neighbours in the image however long their sample chains are.  With one global
// pixel-granular queue they drifted apart — contexts finish one at a time, and every claim landed wherever the
// frame-wide queue head was — and the warp's candidate loads stopped sharing cache lines: at 64 spp the 1 M-sphere scene
// ran at 60 % of its 1-spp rate (profiles/README.md "tile stock").
template <int NC>
__device__ __forceinline__ void body_sample(const Ctx<NC> c, const bool have, const RenderLaunch &p, const float inv_ns, const unsigned lt,
                                            uint32_t &nrays, uint32_t &npaths, uint32_t &stock_next, uint32_t &stock_end, const uint32_t batch) {
    int pix = -1;
    bool need_pixel = false;
    xorwow rng;
    rng.d = rng.v0 = rng.v1 = rng.v2 = rng.v3 = rng.v4 = 0;
    if (have) {
        pix = c.i(F_PIX);
        need_pixel = pix < 0;
        if (!need_pixel) {
            vec3f contrib = mk(0, 0, 0);
            if (c.i(F_HIDX) == -1) {
                const vec3f att = c.v3(F_AX);
                const vec3f k = sky(c.v3(F_DX));
                contrib = mk(mul_(att.x, k.x), mul_(att.y, k.y), mul_(att.z, k.z));
            }
            vec3f col = c.v3(F_CX);
            col = mk(add_(col.x, contrib.x), add_(col.y, contrib.y), add_(col.z, contrib.z));   // main.cu:107
            const uint32_t s = (c.u(F_SD) & 0xffffffu) + 1u;
            if (s >= (uint32_t)p.ns_local) {
                float *out = p.out + (size_t)pix * 3;
                if (p.finalize) {   // main.cu:111-115
                    out[0] = sqrt_(mul_(col.x, inv_ns));
                    out[1] = sqrt_(mul_(col.y, inv_ns));
                    out[2] = sqrt_(mul_(col.z, inv_ns));
                } else {
                    out[0] = col.x; out[1] = col.y; out[2] = col.z;
                }
                need_pixel = true;
            } else {
                c.sv3(F_CX, col);
                c.su(F_SD, s);      // depth 0
                c.load_rng(rng);
            }
        }
    }
    // ---- queue pop for the lanes whose context needs a pixel ----
    const unsigned m = __ballot_sync(0xffffffffu, need_pixel);
    if (m) {
        const uint32_t k = (uint32_t)__popc(m), rank = (uint32_t)__popc(m & lt), avail = stock_end - stock_next;
        uint32_t item = stock_next + rank;
        if (k <= avail) {
            stock_next += k;
        } else {                                   // the stock runs out: the rest comes from a fresh batch of whole tiles
            const uint32_t need = k - avail, claim = (need + batch - 1u) / batch * batch;
            uint32_t qbase = 0;
            const int leader = __ffs(m) - 1;
            if ((int)(threadIdx.x & 31u) == leader) qbase = atomicAdd(p.work_counter, claim);
            qbase = __shfl_sync(0xffffffffu, qbase, leader);
            if (rank >= avail) item = qbase + (rank - avail);
            stock_next = qbase + need;
            stock_end = qbase + claim;
        }
        if (need_pixel) {
            int pi, pj;
            pix = -1;
            if (item >= p.total_items) {
                c.su(F_CW0, 0u); c.su(F_CW1, 0u);                     // S_DONE
                return;
            }
            if (!item_to_pixel(p, item, pi, pj)) {                   // a tile position outside the image: pop again next time
                c.su(F_PIX, (uint32_t)-1);
                c.set_state(S_SAMPLE);
                return;
            }
            pix = pj * p.nx + pi;
            c.su(F_PIX, (uint32_t)pix);
            c.su(F_SD, 0u);
            c.sv3(F_CX, mk(0, 0, 0));
            pixel_stream(p, pix, rng);
        }
    }
    if (have && pix >= 0) {
        gen_sample(c, p, pix, rng, nrays, npaths);
        c.store_rng(rng);
    }
}

template <int NC, int MINB>
__global__ void __launch_bounds__(kRenderThreads, MINB) k_render_pool(const __grid_constant__ RenderLaunch p) {
    extern __shared__ __align__(16) uint32_t pool_smem[];
    constexpr int C = NC / 32;
    constexpr int kWarpWords = NC * NF + 32;
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t *W = pool_smem + (threadIdx.x >> 5) * kWarpWords;
    uint32_t *stage = W + NC * NF;
    const float *planes = &p.tree.planes[0][0];
    const float inv_ns = __fdiv_rn(1.0f, (float)p.ns_total);   // vec3.h:137-144: k = 1.0/t

#pragma unroll
    for (int j = 0; j < C; j++) {       // every context starts in SAMPLE without a pixel: its first visit pops one
        const Ctx<NC> c0{W + lane + 32 * j};
        c0.su(F_PIX, (uint32_t)-1);
        c0.set_state(S_SAMPLE);
    }
    __syncwarp();

    uint32_t nrays = 0, npaths = 0;
    TraceCounters tc;
    tc.sphere_tests = tc.node_tests = tc.voxel_steps = 0;
    uint32_t stock_next = 0, stock_end = 0;      // this warp's unclaimed queue items (body_sample)
    // pixels per context decide the claim size: two tiles with plenty of work, one tile in between, and single pixels
    // (no stock: the old frame-wide queue) when a context sees fewer than 4 pixels — there the tail of the frame is
    // what matters and stocked pixels would only start later (measured, profiles/README.md "tile stock")
    const uint32_t per_ctx = p.total_items / (gridDim.x * (kRenderThreads / 32) * NC);
    const uint32_t batch = per_ctx >= 24u ? 64u : (per_ctx >= 4u ? 32u : 1u);

    uint32_t rounds = 0;
#ifdef RT_COUNTERS
    uint32_t sched_rounds[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, sched_ctx[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
    while (true) {
        // ---- 1. contexts per state: each context stores its own contribution to two words of packed byte counters ----
        uint32_t c0[C], c1[C];
        uint32_t w0 = 0, w1 = 0;
#pragma unroll
        for (int j = 0; j < C; j++) {
            c0[j] = W[F_CW0 * NC + lane + 32 * j];
            c1[j] = W[F_CW1 * NC + lane + 32 * j];
            w0 += c0[j];
            w1 += c1[j];
        }
        w0 = __reduce_add_sync(full, w0);
        w1 = __reduce_add_sync(full, w1);
        if (!(w0 | w1)) break;

        // ---- 2. the state with the most waiting contexts (ties: the earlier, cheaper state): lane s holds the count
        //         of state s, one warp-wide max of (count << 4 | 15 - s) ----
        int best;
        {
            const uint32_t word = lane < 4u ? w0 : w1;
            uint32_t cnt = (word >> ((lane & 3u) * 8u)) & 255u;
            // a TEST round with few contexts wastes the most lanes of the most expensive body: below tune_test_min waiting
            // contexts its count is quartered, so that fuller states (ENTER, STEP, CAND: the ones that feed TEST) run first
            if (lane == (unsigned)S_TEST && cnt < (uint32_t)p.tune_test_min) cnt = (cnt + 3u) >> 2;
            const uint32_t key = lane < 8u ? cnt << 4 | (15u - lane) : 0u;
            best = 15 - (int)(__reduce_max_sync(full, key) & 15u);
        }
        if (++rounds > p.max_rounds) {          // watchdog: a scheduling bug must never hang the GPU; the host reports it
            if (lane == 0) {
                atomicAdd(p.counters + 7, 1ull);
                p.counters[5] = (unsigned long long)w0 << 32 | w1;
                p.counters[6] = (unsigned long long)best << 32 | rounds;
            }
            break;
        }

        // ---- 3. hand up to 32 contexts of that state to the lanes ----
        int id = -1;
        {
            const uint32_t sh = (uint32_t)(best & 3) * 8u;
            const bool hi = best >= 4;
            uint32_t base = 0;
#pragma unroll
            for (int j = 0; j < C; j++) {
                const bool is = (((hi ? c1[j] : c0[j]) >> sh) & 1u) != 0u;
                const unsigned m = __ballot_sync(full, is);
                const uint32_t r = base + (uint32_t)__popc(m & lt);
                if (is && r < 32u) stage[r] = lane + 32u * j;
                base += (uint32_t)__popc(m);
            }
            __syncwarp();
            if (lane < min(base, 32u)) id = (int)stage[lane];
#ifdef RT_COUNTERS
#pragma unroll
            for (int s = 0; s < 10; s++)
                if (s == best) { sched_rounds[s]++; sched_ctx[s] += min(base, 32u); }
#endif
        }
        const Ctx<NC> c{W + (id < 0 ? 0 : id)};

        // ---- 4. the one body of that state (most frequent first) ----
        if (best == S_TEST) body_test<NC>(c, id >= 0, p, tc);
        else if (best == S_CAND) { if (id >= 0) body_cand<NC>(c, p, tc); }
        else if (best == S_STEP) { if (id >= 0) body_step<NC>(c, p, tc); }
        else if (best == S_ENTER) { if (id >= 0) body_enter<NC>(c, p, tc); }
        else if (best == S_END) { if (id >= 0) body_end<NC>(c, p, planes, tc); }
        else if (best == S_SAMPLE) body_sample<NC>(c, id >= 0, p, inv_ns, lt, nrays, npaths, stock_next, stock_end, batch);
        else { if (id >= 0) body_shade<NC>(c, p, nrays); }       // S_DIFF, S_DIEL
        __syncwarp();
    }

    // ---- ray / path counters: one atomic per warp ----
    unsigned long long r64 = nrays, p64 = npaths;
    for (int off = 16; off > 0; off >>= 1) {
        r64 += __shfl_xor_sync(full, r64, off);
        p64 += __shfl_xor_sync(full, p64, off);
    }
    if (lane == 0) {
        atomicAdd(p.counters + 0, r64);
        atomicAdd(p.counters + 1, p64);
    }
#ifdef RT_COUNTERS
    unsigned long long c64[3] = {tc.sphere_tests, tc.node_tests, tc.voxel_steps};
    for (int k = 0; k < 3; k++) {
        for (int off = 16; off > 0; off >>= 1) c64[k] += __shfl_xor_sync(full, c64[k], off);
        if (lane == 0) atomicAdd(p.counters + 2 + k, c64[k]);
    }
    if (lane == 0)
        for (int s = 0; s < 10; s++) {
            atomicAdd(p.counters + 8 + s, (unsigned long long)sched_rounds[s]);
            atomicAdd(p.counters + 18 + s, (unsigned long long)sched_ctx[s]);
        }
#endif
}

template <int NC, int MINB>
static cudaError_t launch_pool(const RenderLaunch &p, int sm_count, cudaStream_t st, int *blocks_out) {
    auto kern = k_render_pool<NC, MINB>;
    const size_t smem = (size_t)(kRenderThreads / 32) * (NC * NF + 32) * sizeof(uint32_t);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRenderThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    long long blocks = (long long)per_sm * sm_count;
    const long long need = ((long long)p.total_items + (long long)(kRenderThreads / 32) * NC - 1) / ((long long)(kRenderThreads / 32) * NC);
    if (blocks > need) blocks = need < 1 ? 1 : need;
    e = cudaMemsetAsync(p.work_counter, 0, 4, st);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)blocks, kRenderThreads, smem, st>>>(p);
    if (blocks_out) *blocks_out = (int)blocks;
    return cudaGetLastError();
}

}  // namespace pool
