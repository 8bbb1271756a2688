// warp_model.cpp — ANALYSIS TOOL (not product, not test): a host model of how a 32-lane warp of k_render spends its loop
// trips, built on the product's own __host__ __device__ traversal code (via tests/hostsim).  It replays the pixel chains of
// whole 8x4 tiles in lock-step exactly as k_render schedules them and records, per closest-hit query, the "script" of the
// flat walk loop (advance / candidate-pair trips, filter positives, exact tests).  From the scripts it derives the lane
// occupancy of the present loop and of alternative organisations, so that kernel designs can be compared before any GPU
// minute is spent.  Driver: profiles/warp_model.py.
#include "../tests/hostsim/hostsim.cpp"

#include <stdio.h>

namespace {

struct RayScript {
    // one entry per visited voxel: number of references in it
    std::vector<uint16_t> vox_counts;
    int positives = 0;       // candidates with disc > 0
    int filter_pass = 0;     // candidates the pre-filter lets through (exact test needed)
    int exact_accept = 0;    // exact tests that improved the closest hit
    int cands_lvl = 0;       // candidates a distance-levelled list would offer (see level_of_ref)
    int pass_cap_hi = 0, pass_cap_both = 0;   // pre-filter passes when roots beyond the voxel's exit (and certain hits before its entry) are left to the other voxels
    int cands_front = 0;     // candidates left when primary rays skip the spheres whose surface inside the voxel faces away from the camera
    int nprolog = 0;
    bool grid_missed = false;
    // per loop trip of the present flat loop: bit0 = advance executed, bits1-2 = candidates tested (0..2), bits3-4 = exact tests run
    std::vector<uint8_t> trips;
};

// How far from the sphere's surface do the hits lie that the FLOAT test reports?  (what rt_build.cuh's sphere_pad must cover)
struct HitError {
    double max_e = 0, max_rel = 0, max_e_m = 0;   // distance from the surface; the same in units of 2^-24 |oc|^2 / r; |oc| of the worst
    double n = 0;
    double hist[40] = {0};                        // bin k: 2^(k-34) <= e < 2^(k-33)
    double bm_max[8] = {0}, bm_n[8] = {0}, bm_p999[8][40] = {{0}};   // by |oc| bucket: [0,1) [1,2) [2,4) [4,8) [8,16) [16,32) [32,64) 64+
    void add(const float4 s, const vec3f o, const vec3f d, const float t) {
        const double px = (double)o.x + (double)t * d.x - s.x, py = (double)o.y + (double)t * d.y - s.y, pz = (double)o.z + (double)t * d.z - s.z;
        const double e = fabs(sqrt(px * px + py * py + pz * pz) - (double)s.w);
        const double ox = (double)o.x - s.x, oy = (double)o.y - s.y, oz = (double)o.z - s.z;
        const double m2 = ox * ox + oy * oy + oz * oz;
        const double rel = e * s.w / m2 * 16777216.0;
        if (e > max_e) { max_e = e; max_e_m = sqrt(m2); }
        if (rel > max_rel) max_rel = rel;
        int k = e > 0 ? (int)floor(log2(e)) + 34 : 0;
        k = k < 0 ? 0 : (k > 39 ? 39 : k);
        hist[k]++;
        n++;
        const double m = sqrt(m2);
        const int b = m < 1 ? 0 : (m < 2 ? 1 : (m < 4 ? 2 : (m < 8 ? 3 : (m < 16 ? 4 : (m < 32 ? 5 : (m < 64 ? 6 : 7))))));
        bm_n[b]++; bm_p999[b][k]++;
        if (e > bm_max[b]) bm_max[b] = e;
    }
};
static thread_local HitError t_err;
static HitError g_err;

// distance levels of a voxel list: reference k of a voxel is needed only by rays that have travelled far enough for the float error of
// the hit test (~ K |oc|^2 / r) to reach across the gap between the sphere's surface and the voxel
static const float kLvlDist[4] = {8.0f, 16.0f, 28.0f, 64.0f};
static float g_lvl_k = getenv("WM_LVL_K") ? (float)atof(getenv("WM_LVL_K")) : 1e-6f;
static int level_of_ref(const GridView &g, const float4 s, int ix, int iy, int iz) {
    float lo[3], hi[3];
    voxel_box(g, ix, iy, iz, lo, hi);
    for (int l = 0; l < 3; l++) {
        const float m = kLvlDist[l] + s.w;
        if (shell_hits_box(s, 2e-4f + g_lvl_k * m * m / fmaxf(s.w, 1e-3f), lo, hi)) return l;
    }
    return 3;
}

// Thales: p on sphere (c, r) faces the eye e, (p - e).(p - c) <= 0, iff p lies inside the sphere on the diameter e-c.  A voxel can
// hold a front-facing point of the surface only if it reaches into that sphere (margin: lens radius + list padding).
static bool g_primary = false;
#pragma omp threadprivate(g_primary)
static float g_eye[3] = {0, 0, 0}, g_eye_margin = 0.06f;
static bool faces_eye(const GridView &g, const float4 s, int ix, int iy, int iz) {
    float lo[3], hi[3];
    voxel_box(g, ix, iy, iz, lo, hi);
    const float m[3] = {0.5f * (g_eye[0] + s.x), 0.5f * (g_eye[1] + s.y), 0.5f * (g_eye[2] + s.z)};
    const float ex = g_eye[0] - s.x, ey = g_eye[1] - s.y, ez = g_eye[2] - s.z;
    const float R2 = 0.25f * (ex * ex + ey * ey + ez * ez) + (s.w + sphere_pad(s.w)) * g_eye_margin;   // (p - e).(p - c) <= |p - c| * margin
    float d2 = 0;
    for (int k = 0; k < 3; k++) {
        const float a = lo[k] - m[k], b = hi[k] - m[k];
        const float nearest = a > 0.f ? a : (b < 0.f ? b : 0.f);
        d2 += nearest * nearest;
    }
    return d2 <= R2;
}

// trace_walk<false> (rt_trace.cuh) with recording; `two` = candidates per trip
static Hit walk_record(const SceneView &sc, const TreeView &tv, const vec3f o, const vec3f d, RayScript &rs, const int per_trip) {
    Hit h;
    h.t = kTMax; h.idx = -1;
    RayPre r;
    r.o = o; r.d = d;
    r.a = dot3(d, d);
    const float ia = rcp_trav(r.a);
    { float t; if (sphere_test(sc.geom[0], o, d, r.a, kTMax, t)) { h.t = t; h.idx = 0; } }
    rs.nprolog = tv.nprolog;
    for (int k = 1; k < tv.nprolog; k++) {
        const int idx = (int)tv.prolog[k];
        float t;
        const float4 s = sc.geom[idx];
        if (maybe_hit(s, o, d, r.a, ia, h.t) && sphere_test(s, o, d, r.a, h.t, t)) { h.t = t; h.idx = idx; }
    }
    const GridView &g = tv.grid;
    if (g.nx == 0) return h;
    r.inv = mk(rcp_trav(d.x), rcp_trav(d.y), rcp_trav(d.z));
    float te, tx;
    if (!ray_box(r, g.org, g.hi, h.t * (1.0f + kTSlackRel) + kTSlackAbs, te, tx)) { rs.grid_missed = true; return h; }
    int ix = (int)floorf((o.x + d.x * te - g.org[0]) * g.inv_vs[0]);
    int iy = (int)floorf((o.y + d.y * te - g.org[1]) * g.inv_vs[1]);
    int iz = (int)floorf((o.z + d.z * te - g.org[2]) * g.inv_vs[2]);
    ix = imin(imax(ix, 0), g.nx - 1);
    iy = imin(imax(iy, 0), g.ny - 1);
    iz = imin(imax(iz, 0), g.nz - 1);
    const int sx = d.x >= 0.0f ? 1 : -1, sy = d.y >= 0.0f ? 1 : -1, sz = d.z >= 0.0f ? 1 : -1;
    float tmx = fabsf(d.x) > 0.0f ? (g.org[0] + (float)(ix + (sx > 0)) * g.vs[0] - o.x) * r.inv.x : kTMax;
    float tmy = fabsf(d.y) > 0.0f ? (g.org[1] + (float)(iy + (sy > 0)) * g.vs[1] - o.y) * r.inv.y : kTMax;
    float tmz = fabsf(d.z) > 0.0f ? (g.org[2] + (float)(iz + (sz > 0)) * g.vs[2] - o.z) * r.inv.z : kTMax;
    const float dtx = fabsf(g.vs[0] * r.inv.x), dty = fabsf(g.vs[1] * r.inv.y), dtz = fabsf(g.vs[2] * r.inv.z);
    uint32_t k, e;
    float t_vox_in = te;
    { const uint2 v = g.vox[((size_t)(iz * g.ny + iy) * g.nx + ix)]; k = v.x; e = v.x + v.y; rs.vox_counts.push_back((uint16_t)v.y); }
    int budget = g.nx + g.ny + g.nz + 4;
    bool walking = true;
    while (walking) {
        uint8_t trip = 0;
        if (k >= e) {
            trip |= 1;
            float t_in;
            if (tmx <= tmy && tmx <= tmz) { t_in = tmx; ix += sx; tmx += dtx; walking = (unsigned)ix < (unsigned)g.nx; }
            else if (tmy <= tmz)          { t_in = tmy; iy += sy; tmy += dty; walking = (unsigned)iy < (unsigned)g.ny; }
            else                          { t_in = tmz; iz += sz; tmz += dtz; walking = (unsigned)iz < (unsigned)g.nz; }
            if (t_in > h.t * (1.0f + kTSlackRel) + kTSlackAbs || t_in > tx * (1.0f + 1e-5f) + 1e-6f || --budget < 0) walking = false;
            if (walking) {
                const uint2 v = g.vox[((size_t)(iz * g.ny + iy) * g.nx + ix)];
                k = v.x; e = v.x + v.y;
                rs.vox_counts.push_back((uint16_t)v.y);
                t_vox_in = t_in;
            }
        }
        if (walking && k < e) {
            int ntest = 0, nexact = 0;
            const float bound = h.t;
            const float l_out = fminf(tmx, fminf(tmy, tmz)) * sqrtf(r.a) * 1.001f;
            const int ray_lvl = (l_out > kLvlDist[0]) + (l_out > kLvlDist[1]) + (l_out > kLvlDist[2]);
            for (int q = 0; q < per_trip && k < e; q++, k++) {
                const float4 s = sc.geom[g.refs[k]];
                ntest++;
                if (level_of_ref(g, s, ix, iy, iz) <= ray_lvl) rs.cands_lvl++;
                if (!g_primary || faces_eye(g, s, ix, iy, iz)) rs.cands_front++;
                const vec3f oc = mk(sub_(o.x, s.x), sub_(o.y, s.y), sub_(o.z, s.z));
                const float b = dot3(oc, d), c = fma_(-s.w, s.w, dot3(oc, oc)), disc = fma_(b, b, -mul_(r.a, c));
                if (disc > 0.0f) {
                    rs.positives++;
                    float ta;
                    if (s.w < 0.5f && sphere_test(s, o, d, r.a, kTMax, ta)) t_err.add(s, o, d, ta);     // (small spheres: the gridded ones)
                }
                {
                    const float t_out_v = fminf(tmx, fminf(tmy, tmz));
                    const float cap = fminf(bound, t_out_v * (1.0f + 1e-4f) + 1e-4f);
                    float ubx;
                    if (maybe_hit_ub(s, o, d, r.a, ia, cap, ubx)) {
                        rs.pass_cap_hi++;
                        // certain hit that lies before this voxel: the voxel holding it has offered it already
                        const float sa = sqrtf(fmaxf(disc, 0.0f)), t1 = (-b - sa) * ia, eps = (fabsf(b) + sa) * ia * 1e-5f;
                        if (!(t1 - eps > kTMin && t1 + eps < t_vox_in * (1.0f - 1e-4f) - 1e-4f)) rs.pass_cap_both++;
                    }
                }
                if (maybe_hit(s, o, d, r.a, ia, bound)) {
                    rs.filter_pass++;
                    nexact++;
                    float t;
                    if (sphere_test(s, o, d, r.a, h.t, t)) { h.t = t; h.idx = (int)g.refs[k]; rs.exact_accept++; }
                }
            }
            trip |= (uint8_t)(ntest << 1) | (uint8_t)(nexact << 3);
        }
        rs.trips.push_back(trip);
    }
    return h;
}

struct Lane {
    int pix = -1, pi = 0, pj = 0, s = 0, depth = 0;
    xorwow rng;
    vec3f o, d;
};

struct Acc {
    // per-ray totals
    double rays = 0, paths = 0, vox_visits = 0, vox_nonempty = 0, cands = 0, positives = 0, filter_pass = 0, exact_accept = 0, grid_missed = 0;
    double trips = 0, cands_lvl = 0, cands_front = 0, pass_cap_hi = 0, pass_cap_both = 0;
    // present loop, lock-step
    double outer = 0, active_lane_outer = 0;         // outer iterations (one closest-hit query per active lane), lanes with a pixel
    double loop_trips = 0;                           // warp loop trips (max over lanes)
    double trips_with_adv = 0, lanes_adv = 0, trips_with_test = 0, lanes_test = 0, trips_with_exact = 0, lanes_exact = 0;
    // design B: voxel rounds with flattened candidates
    double b_rounds = 0, b_chunks = 0, b_lanes_round = 0;
    // design B2: rounds advance to the next NON-EMPTY voxel (empties skipped inside the advance)
    double b2_rounds = 0, b2_chunks = 0, b2_max_adv = 0;
    // what k_render_coop does: B2 with ITEMS = 2 / 4 consecutive references per lane and chunk step
    double c2_chunks = 0, c2_slots_used = 0, c4_chunks = 0, c4_slots_used = 0;
    // variant "deferred tails": a round runs only its FULL chunks; the items left over stay with their (trailing) lanes and are
    // republished next round next to the other lanes' new voxels (a partial chunk runs only when there is no full one)
    double d2_chunks = 0, d2_rounds = 0, d4_chunks = 0, d4_rounds = 0;
    // histograms
    double hist_vox[65] = {0}, hist_cands[257] = {0}, hist_trips[129] = {0};
    double shade_hit = 0, shade_sky = 0;
    double mat[3] = {0, 0, 0};
};

}  // namespace

extern "C" int wm_run(const hs_sphere *sph, int n, const float *camera22, const void *blob, int spl, float density, int nx, int ny, int ns,
                      int max_depth, int tile_first, int tile_step, int warps, int tiles_per_warp, int per_trip, double *out, int out_len) {
    std::vector<float4> geom((size_t)n), matl((size_t)n);
    std::vector<int> tag((size_t)n);
    for (int i = 0; i < n; i++) {
        geom[(size_t)i] = make_float4(sph[i].cx, sph[i].cy, sph[i].cz, sph[i].radius);
        matl[(size_t)i] = make_float4(sph[i].ax, sph[i].ay, sph[i].az, sph[i].param);
        tag[(size_t)i] = sph[i].mat;
    }
    SceneView sc;
    sc.geom = geom.data(); sc.matl = matl.data(); sc.tag = tag.data(); sc.n = n;
    HostTree T;
    TreeView tv;
    build_host_tree(geom, tag, static_cast<const int32_t *>(blob), spl, density, T);
    view_of(T, tv);
    fprintf(stderr, "grid %d x %d x %d, voxels %zu refs %zu, voxel size %.4f %.4f %.4f, prolog %d\n", tv.grid.nx, tv.grid.ny, tv.grid.nz, T.vox.size(),
            T.refs.size(), tv.grid.vs[0], tv.grid.vs[1], tv.grid.vs[2], tv.nprolog);
    CameraData cam;
    memcpy(&cam, camera22, sizeof cam);
    const int tiles_x = (nx + 7) / 8, tiles_y = (ny + 3) / 4;
    for (int k = 0; k < 3; k++) g_eye[k] = cam.origin[k];
    g_eye_margin = cam.lens_radius + 0.01f;
    Acc A;
#pragma omp parallel
    {
        Acc a;
#pragma omp for schedule(dynamic, 1)
        for (int w = 0; w < warps; w++) {
            Lane L[32];
            uint32_t stock_next = 0, stock_end = 0;
            int tiles_taken = 0;
            const int base_tile = tile_first + w * tile_step;
            while (true) {
                // claim
                for (int l = 0; l < 32; l++) {
                    while (L[l].pix < 0) {
                        if (stock_next >= stock_end) {
                            if (tiles_taken >= tiles_per_warp) break;
                            stock_next = (uint32_t)(base_tile + tiles_taken) * 32u;
                            stock_end = stock_next + 32u;
                            tiles_taken++;
                        }
                        const uint32_t item = stock_next++;
                        const uint32_t tile = item >> 5, in = item & 31u;
                        if ((int)tile >= tiles_x * tiles_y) continue;
                        const int ty = (int)tile / tiles_x, tx_ = (int)tile - ty * tiles_x;
                        const int i = tx_ * 8 + (int)(in & 7u), j = ty * 4 + (int)(in >> 3);
                        if (i >= nx || j >= ny) continue;
                        L[l].pi = i; L[l].pj = j; L[l].pix = j * nx + i; L[l].s = 0; L[l].depth = 0;
                        xorwow_seed(L[l].rng, (unsigned long long)(long long)(1984 + L[l].pix));
                    }
                }
                int nact = 0;
                for (int l = 0; l < 32; l++) nact += L[l].pix >= 0;
                if (!nact) break;
                a.outer++;
                a.active_lane_outer += nact;
                RayScript rs[32];
                vec3f att_dummy = mk(1, 1, 1);
                (void)att_dummy;
                Hit hs[32];
                for (int l = 0; l < 32; l++) {
                    if (L[l].pix < 0) continue;
                    Lane &q = L[l];
                    if (q.depth == 0) {
                        const float u = div_(add_((float)q.pi, xorwow_uniform(q.rng)), (float)nx);
                        const float v = div_(add_((float)q.pj, xorwow_uniform(q.rng)), (float)ny);
                        camera_ray(cam, u, v, q.rng, q.o, q.d);
                        a.paths++;
                    }
                    a.rays++;
                    TraceCounters tcn{};
                    hs[l] = trace_tree(sc, tv, &tv.planes[0][0], q.o, q.d, tcn);
                    g_primary = q.depth == 0;
                    const Hit h2 = walk_record(sc, tv, q.o, q.d, rs[l], per_trip);
                    if (h2.idx != hs[l].idx || h2.t != hs[l].t) { /* checked re-walk case: rare; keep trace_tree's answer */ }
                    const RayScript &r = rs[l];
                    a.vox_visits += r.vox_counts.size();
                    int c = 0;
                    for (uint16_t vc : r.vox_counts) { a.vox_nonempty += vc > 0; c += vc; }
                    a.cands += c; a.positives += r.positives; a.filter_pass += r.filter_pass; a.exact_accept += r.exact_accept;
                    a.cands_lvl += r.cands_lvl; a.cands_front += r.cands_front; a.pass_cap_hi += r.pass_cap_hi; a.pass_cap_both += r.pass_cap_both;
                    a.grid_missed += r.grid_missed;
                    a.trips += r.trips.size();
                    a.hist_vox[std::min<size_t>(r.vox_counts.size(), 64)]++;
                    a.hist_cands[std::min(c, 256)]++;
                    a.hist_trips[std::min<size_t>(r.trips.size(), 128)]++;
                }
                // ---- present loop in lock-step ----
                size_t tmax = 0;
                for (int l = 0; l < 32; l++) tmax = std::max(tmax, rs[l].trips.size());
                a.loop_trips += tmax;
                for (size_t t = 0; t < tmax; t++) {
                    int nadv = 0, ntest = 0, nex = 0;
                    for (int l = 0; l < 32; l++) {
                        if (t >= rs[l].trips.size()) continue;
                        const uint8_t b = rs[l].trips[t];
                        nadv += b & 1; ntest += ((b >> 1) & 3) > 0; nex += ((b >> 3) & 3) > 0;
                    }
                    a.trips_with_adv += nadv > 0; a.lanes_adv += nadv;
                    a.trips_with_test += ntest > 0; a.lanes_test += ntest;
                    a.trips_with_exact += nex > 0; a.lanes_exact += nex;
                }
                // ---- design B: one voxel per lane per round, candidates flattened over the warp ----
                size_t vmax = 0;
                for (int l = 0; l < 32; l++) vmax = std::max(vmax, rs[l].vox_counts.size());
                a.b_rounds += vmax;
                for (size_t t = 0; t < vmax; t++) {
                    int tot = 0, lanes = 0;
                    for (int l = 0; l < 32; l++) if (t < rs[l].vox_counts.size()) { tot += rs[l].vox_counts[t]; lanes++; }
                    a.b_chunks += (tot + 31) / 32;
                    a.b_lanes_round += lanes;
                }
                // ---- design B2: a round = next non-empty voxel of every lane ----
                {
                    std::vector<std::vector<std::pair<int, int>>> ne(32);   // (count, empties skipped before it)
                    size_t rmax = 0;
                    for (int l = 0; l < 32; l++) {
                        int skipped = 0;
                        for (uint16_t vc : rs[l].vox_counts) { if (vc) { ne[l].push_back({vc, skipped}); skipped = 0; } else skipped++; }
                        if (skipped) ne[l].push_back({0, skipped});          // trailing empties until the walk ends
                        rmax = std::max(rmax, ne[l].size());
                    }
                    a.b2_rounds += rmax;
                    for (size_t t = 0; t < rmax; t++) {
                        int tot = 0, madv = 0, p2 = 0, p4 = 0;
                        for (int l = 0; l < 32; l++) if (t < ne[l].size()) {
                            tot += ne[l][t].first; madv = std::max(madv, ne[l][t].second + 1);
                            p2 += (ne[l][t].first + 1) / 2; p4 += (ne[l][t].first + 3) / 4;
                        }
                        a.b2_chunks += (tot + 31) / 32;
                        a.b2_max_adv += madv;
                        a.c2_chunks += (p2 + 31) / 32; a.c2_slots_used += tot;
                        a.c4_chunks += (p4 + 31) / 32; a.c4_slots_used += tot;
                    }
                }
                // ---- deferred tails ----
                for (int items_per = 2; items_per <= 4; items_per += 2) {
                    std::vector<std::vector<int>> seq(32);
                    for (int l = 0; l < 32; l++)
                        for (uint16_t vc : rs[l].vox_counts) if (vc) seq[l].push_back((vc + items_per - 1) / items_per);
                    size_t pos[32] = {0};
                    int rem[32];
                    for (int l = 0; l < 32; l++) rem[l] = seq[l].empty() ? 0 : seq[l][0];
                    double chunks = 0, rounds = 0;
                    while (true) {
                        int total = 0;
                        for (int l = 0; l < 32; l++) total += rem[l];
                        if (!total) break;
                        rounds++;
                        int run = total >= 32 ? (total / 32) * 32 : total;
                        chunks += (run + 31) / 32;
                        for (int l = 0; l < 32 && run > 0; l++) {
                            const int take = std::min(run, rem[l]);
                            rem[l] -= take; run -= take;
                        }
                        for (int l = 0; l < 32; l++)
                            if (rem[l] == 0 && pos[l] < seq[l].size()) { pos[l]++; rem[l] = pos[l] < seq[l].size() ? seq[l][pos[l]] : 0; }
                    }
                    if (items_per == 2) { a.d2_chunks += chunks; a.d2_rounds += rounds; } else { a.d4_chunks += chunks; a.d4_rounds += rounds; }
                }
                // ---- shade, exactly as k_render ----
                for (int l = 0; l < 32; l++) {
                    if (L[l].pix < 0) continue;
                    Lane &q = L[l];
                    const Hit h = hs[l];
                    bool sample_done = false;
                    if (h.idx >= 0) {
                        a.shade_hit++;
                        a.mat[std::min(std::max(tag[(size_t)h.idx], 0), 2)]++;
                        vec3f hp, hn, at, dn;
                        hit_point(geom[(size_t)h.idx], q.o, q.d, h.t, hp, hn);
                        if (scatter(tag[(size_t)h.idx], matl[(size_t)h.idx], q.d, hp, hn, at, dn, q.rng)) {
                            q.o = hp; q.d = dn; q.depth++;
                            if (q.depth >= max_depth) sample_done = true;
                        } else sample_done = true;
                    } else { a.shade_sky++; sample_done = true; }
                    if (sample_done) { q.depth = 0; q.s++; if (q.s >= ns) q.pix = -1; }
                }
            }
        }
#pragma omp critical
        {
            g_err.n += t_err.n;
            if (t_err.max_e > g_err.max_e) { g_err.max_e = t_err.max_e; g_err.max_e_m = t_err.max_e_m; }
            g_err.max_rel = std::max(g_err.max_rel, t_err.max_rel);
            for (int k = 0; k < 40; k++) g_err.hist[k] += t_err.hist[k];
            for (int b = 0; b < 8; b++) {
                g_err.bm_n[b] += t_err.bm_n[b];
                g_err.bm_max[b] = std::max(g_err.bm_max[b], t_err.bm_max[b]);
                for (int k = 0; k < 40; k++) g_err.bm_p999[b][k] += t_err.bm_p999[b][k];
            }
            t_err = HitError();
            double *dst = reinterpret_cast<double *>(&A), *src = reinterpret_cast<double *>(&a);
            for (size_t i = 0; i < sizeof(Acc) / sizeof(double); i++) dst[i] += src[i];
        }
    }
    fprintf(stderr, "float-accepted hits on gridded spheres: %.0f; distance from the surface: max %.3g (|oc| %.1f), max in units of 2^-24 |oc|^2 / r: %.2f\n",
            g_err.n, g_err.max_e, g_err.max_e_m, g_err.max_rel);
    fprintf(stderr, "  log2(distance) histogram:");
    for (int k = 0; k < 40; k++) if (g_err.hist[k] > 0) fprintf(stderr, " [%d] %.0f", k - 34, g_err.hist[k]);
    fprintf(stderr, "\n");
    for (int b = 0; b < 8; b++) {
        if (g_err.bm_n[b] == 0) continue;
        double cum = 0; int k999 = 0;
        for (int k = 0; k < 40; k++) { cum += g_err.bm_p999[b][k]; if (cum >= 0.999 * g_err.bm_n[b]) { k999 = k; break; } }
        fprintf(stderr, "  |oc| bucket %d: hits %.0f, max distance %.3g, 99.9 %% below 2^%d\n", b, g_err.bm_n[b], g_err.bm_max[b], k999 - 33);
    }
    g_err = HitError();
    const size_t nd = sizeof(Acc) / sizeof(double);
    if ((size_t)out_len < nd) return -(int)nd;
    memcpy(out, &A, sizeof A);
    return (int)nd;
}
