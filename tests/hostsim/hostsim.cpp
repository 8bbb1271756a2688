// hostsim.cpp — TEST-ONLY host emulation of the product's per-lane device code (never linked into librt_b200.so).
//
// The render kernel's arithmetic lives in __host__ __device__ headers (rt_shade.cuh, rt_trace.cuh, rt_build.cuh)
// whose float primitives map to explicitly rounded operations on both sides (rt_math.cuh).  Compiling them for
// the host lets the CPU-only test tier run the PRODUCT's closest-hit and shading code against the oracle without
// a GPU: it catches traversal / margin / arithmetic-pattern bugs before any GPU minute is spent.  What it cannot
// cover — the build kernels, the persistent scheduling, libdevice powf/tanf — is covered by the -m gpu tests.
//
// The traversal structure is rebuilt here on the host from a reference-layout Octree blob (so the candidate lists
// are exactly the reference's) using the same per-item functions the build kernels call (sphere_pad, choose_grid,
// voxel_range, shell_hits_box).  This is not a fallback path: the library has none.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../dd2360-raytracing_b200/csrc/rt_build.cuh"
#include "../../dd2360-raytracing_b200/csrc/rt_trace.cuh"

using namespace rt;

static inline float4 make_float4(float x, float y, float z, float w) { float4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

struct hs_sphere { float cx, cy, cz, radius; int32_t mat; float ax, ay, az, param; };
struct hs_params { int nx, ny, ns, use_octree, spl, arith, seed_mode, max_depth, i0, i1, istep, j0, j1, jstep, threads; };
struct hs_counters { uint64_t rays, sphere_tests, aabb_tests, paths; uint32_t max_depth; };

struct HostTree {
    std::vector<TreeNode> nodes;
    std::vector<TreeExtent> node_ext;
    std::vector<CellGrid> cells;
    std::vector<TreeExtent> cell_ext;
    std::vector<uint32_t> vox_start, vox_refs, big_refs;
    float planes[3][kPlanes];
};

static void make_planes(float P[3][kPlanes]) {
    const float lo[3] = {-11.f, 0.f, -11.f}, hi[3] = {11.f, 2.f, 11.f};
    for (int a = 0; a < 3; a++) {
        P[a][0] = lo[a]; P[a][8] = hi[a];
        for (int step = 8; step > 1; step >>= 1)
            for (int i = 0; i < 8; i += step) P[a][i + step / 2] = P[a][i] + (P[a][i + step] - P[a][i]) / 2;
    }
}

static void build_host_tree(const std::vector<float4> &geom, const std::vector<int> &tag, const int32_t *blob, int spl,
                            float density, HostTree &T) {
    make_planes(T.planes);
    const int32_t *bn = blob;
    const int32_t *bl = blob + kNumberNodes * kNodeInts;
    const int32_t *cnt = bl + (size_t)(kNumberLeafs + 1) * (size_t)(spl + 1);
    const int node_count = cnt[0];
    T.nodes.assign((size_t)node_count, TreeNode{});
    T.node_ext.assign((size_t)node_count, TreeExtent{});
    for (auto &x : T.node_ext) for (int k = 0; k < 3; k++) { x.lo[k] = 3e38f; x.hi[k] = -3e38f; }
    T.vox_start.clear(); T.vox_start.push_back(0);
    auto plane_index = [&](int a, float v) { for (int i = 0; i < kPlanes; i++) if (T.planes[a][i] == v) return i; return -1; };
    for (int ni = 0; ni < node_count; ni++) {
        const int32_t *nd = bn + ni * kNodeInts;
        const float *bx = reinterpret_cast<const float *>(nd + 1);
        TreeNode &tn = T.nodes[(size_t)ni];
        memset(&tn, 0, sizeof tn);
        tn.level = (uint8_t)nd[0];
        const int sh = 3 - nd[0];
        tn.ix = (uint8_t)(plane_index(0, bx[0]) >> sh);
        tn.iy = (uint8_t)(plane_index(1, bx[1]) >> sh);
        tn.iz = (uint8_t)(plane_index(2, bx[2]) >> sh);
        tn.first_cell = 0xffffffffu;
        if (nd[0] < 3) {
            for (int c = 0; c < 8; c++) tn.child[c] = (uint16_t)nd[7 + c];
            continue;
        }
        // level 3: gather the stored list from the leaf buckets
        std::vector<uint32_t> small, big;
        const float ex = T.planes[0][tn.ix + 1] - T.planes[0][tn.ix], ey = T.planes[1][tn.iy + 1] - T.planes[1][tn.iy],
                    ez = T.planes[2][tn.iz + 1] - T.planes[2][tn.iz];
        const float big_r = kBigRadiusFrac * fmaxf(ex, fmaxf(ey, ez));
        for (int c = 0; c < 8; c++) {
            const int leaf = nd[7 + c];
            if (leaf == 0) break;
            const int32_t *lf = bl + (size_t)leaf * (size_t)(spl + 1);
            for (int j = 0; j < lf[spl]; j++) {
                const uint32_t idx = (uint32_t)lf[j];
                if (tag[idx] < 0) continue;
                if (geom[idx].w > big_r && (int)big.size() < kMaxBigPerCell) big.push_back(idx);
                else small.push_back(idx);
            }
        }
        if (small.empty() && big.empty()) continue;
        CellGrid g;
        memset(&g, 0, sizeof g);
        g.morton = (uint32_t)morton_of(tn.ix, tn.iy, tn.iz);
        float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
        for (uint32_t idx : small) {
            const float4 s = geom[idx];
            const float r = s.w + sphere_pad(s.w);
            lo[0] = fminf(lo[0], s.x - r); hi[0] = fmaxf(hi[0], s.x + r);
            lo[1] = fminf(lo[1], s.y - r); hi[1] = fmaxf(hi[1], s.y + r);
            lo[2] = fminf(lo[2], s.z - r); hi[2] = fmaxf(hi[2], s.z + r);
        }
        const uint32_t voxels = choose_grid(lo, hi, (uint32_t)small.size(), density, g);
        g.vox_base = (uint32_t)T.vox_start.size() - 1;
        const int dense = (int)T.cells.size();
        g.big = ((uint32_t)dense * kMaxBigPerCell) << 8 | (uint32_t)big.size();
        T.big_refs.resize((size_t)(dense + 1) * kMaxBigPerCell, 0);
        std::sort(big.begin(), big.end());
        for (size_t a = 0; a < big.size(); a++) T.big_refs[(size_t)dense * kMaxBigPerCell + a] = big[a];
        if (voxels) {
            std::vector<std::vector<uint32_t>> lists(voxels);
            const int gx = (int)(g.dims & 1023u), gy = (int)((g.dims >> 10) & 1023u);
            for (uint32_t idx : small) {
                const float4 s = geom[idx];
                const float pad = sphere_pad(s.w);
                int v0[3], v1[3];
                voxel_range(g, s, pad, v0, v1);
                for (int z = v0[2]; z <= v1[2]; z++)
                    for (int y = v0[1]; y <= v1[1]; y++)
                        for (int x = v0[0]; x <= v1[0]; x++) {
                            float blo[3], bhi[3];
                            voxel_box(g, x, y, z, blo, bhi);
                            if (shell_hits_box(s, pad, blo, bhi)) lists[(size_t)((z * gy + y) * gx + x)].push_back(idx);
                        }
            }
            for (auto &l : lists) {
                std::sort(l.begin(), l.end());
                T.vox_refs.insert(T.vox_refs.end(), l.begin(), l.end());
                T.vox_start.push_back((uint32_t)T.vox_refs.size());
            }
        }
        TreeExtent xe;
        for (int k = 0; k < 3; k++) { xe.lo[k] = voxels ? g.org[k] : 3e38f; xe.hi[k] = voxels ? g.hi[k] : -3e38f; }
        for (uint32_t idx : big) {
            const float4 s = geom[idx];
            const float r = s.w + sphere_pad(s.w);
            xe.lo[0] = fminf(xe.lo[0], s.x - r); xe.hi[0] = fmaxf(xe.hi[0], s.x + r);
            xe.lo[1] = fminf(xe.lo[1], s.y - r); xe.hi[1] = fmaxf(xe.hi[1], s.y + r);
            xe.lo[2] = fminf(xe.lo[2], s.z - r); xe.hi[2] = fmaxf(xe.hi[2], s.z + r);
        }
        xe.pad[0] = xe.pad[1] = 0;
        tn.first_cell = (uint32_t)dense;
        T.cells.push_back(g);
        T.cell_ext.push_back(xe);
        T.node_ext[(size_t)ni] = xe;
    }
    // extents bottom-up (children always have larger indices than their parent in creation order? not
    // guaranteed across subtrees, so iterate by level)
    for (int lv = 2; lv >= 0; lv--)
        for (int ni = 0; ni < node_count; ni++) {
            if (T.nodes[(size_t)ni].level != lv) continue;
            TreeExtent &x = T.node_ext[(size_t)ni];
            for (int c = 0; c < 8; c++) {
                const int ch = T.nodes[(size_t)ni].child[c];
                if (!ch) continue;
                for (int k = 0; k < 3; k++) {
                    x.lo[k] = fminf(x.lo[k], T.node_ext[(size_t)ch].lo[k]);
                    x.hi[k] = fmaxf(x.hi[k], T.node_ext[(size_t)ch].hi[k]);
                }
            }
        }
}

extern "C" {

// camera22: origin, llc, horizontal, vertical, u, v, w, lens_radius (taken from the oracle so that only the
// hot path is under test here)
static int render_core(const hs_sphere *sph, int n, const float *camera22, const void *blob, const hs_params *p, float density,
              float *fb_gamma, float *fb_linear, hs_counters *ctr_out, uint64_t *build_stats /* voxels, refs */,
              const TreeView *given) {
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
    memset(&tv, 0, sizeof tv);
    if (p->use_octree && given) {
        tv = *given;
        make_planes(T.planes);
    } else if (p->use_octree) {
        build_host_tree(geom, tag, static_cast<const int32_t *>(blob), p->spl, density, T);
        tv.nodes = T.nodes.data(); tv.node_ext = T.node_ext.data(); tv.cells = T.cells.data();
        tv.cell_ext = T.cell_ext.data(); tv.vox_start = T.vox_start.data(); tv.vox_refs = T.vox_refs.data();
        tv.big_refs = T.big_refs.data();
        tv.node_count = (int)T.nodes.size(); tv.cell_count = (int)T.cells.size();
        if (build_stats) { build_stats[0] = T.vox_start.size() - 1; build_stats[1] = T.vox_refs.size(); }
    }
    CameraData cam;
    memcpy(&cam, camera22, sizeof cam);
    hs_counters total;
    memset(&total, 0, sizeof total);
    const int nrows = (p->j1 - p->j0 + p->jstep - 1) / p->jstep;
#pragma omp parallel
    {
        hs_counters c;
        memset(&c, 0, sizeof c);
#pragma omp for schedule(dynamic, 1)
        for (int r = 0; r < nrows; r++) {
            const int j = p->j0 + r * p->jstep;
            for (int i = p->i0; i < p->i1; i += p->istep) {
                const int pix = j * p->nx + i;
                xorwow rng;
                xorwow_seed(rng, (unsigned long long)(long long)(1984 + pix));
                vec3f col = mk(0, 0, 0);
                for (int s = 0; s < p->ns; s++) {
                    const float u = div_(add_((float)i, xorwow_uniform(rng)), (float)p->nx);
                    const float v = div_(add_((float)j, xorwow_uniform(rng)), (float)p->ny);
                    vec3f o, d, att = mk(1, 1, 1), contrib = mk(0, 0, 0);
                    camera_ray(cam, u, v, rng, o, d);
                    c.paths++;
                    for (int depth = 0; depth < p->max_depth; depth++) {
                        c.rays++;
                        TraceCounters tcn{};
                        Hit h = p->use_octree ? trace_tree(sc, tv, &T.planes[0][0], o, d, tcn)
                                              : trace_list(sc.geom, sc.tag, sc.n, o, d, tcn);
                        if (h.idx >= 0) {
                            vec3f hp, hn, a, dn;
                            hit_point(geom[(size_t)h.idx], o, d, h.t, hp, hn);
                            if (!scatter(tag[(size_t)h.idx], matl[(size_t)h.idx], d, hp, hn, a, dn, rng)) break;
                            att = mk(mul_(att.x, a.x), mul_(att.y, a.y), mul_(att.z, a.z));
                            o = hp; d = dn;
                        } else {
                            const vec3f k = sky(d);
                            contrib = mk(mul_(att.x, k.x), mul_(att.y, k.y), mul_(att.z, k.z));
                            break;
                        }
                    }
                    col = mk(add_(col.x, contrib.x), add_(col.y, contrib.y), add_(col.z, contrib.z));
                }
                if (fb_linear) { fb_linear[3 * (size_t)pix] = col.x; fb_linear[3 * (size_t)pix + 1] = col.y; fb_linear[3 * (size_t)pix + 2] = col.z; }
                if (fb_gamma) {
                    const float k = div_(1.0f, (float)p->ns);
                    fb_gamma[3 * (size_t)pix] = sqrt_(mul_(col.x, k));
                    fb_gamma[3 * (size_t)pix + 1] = sqrt_(mul_(col.y, k));
                    fb_gamma[3 * (size_t)pix + 2] = sqrt_(mul_(col.z, k));
                }
            }
        }
#pragma omp critical
        { total.rays += c.rays; total.paths += c.paths; }
    }
    if (ctr_out) *ctr_out = total;
    return 0;
}


int hs_render(const hs_sphere *sph, int n, const float *camera22, const void *blob, const hs_params *p, float density,
              float *fb_gamma, float *fb_linear, hs_counters *ctr_out, uint64_t *build_stats) {
    return render_core(sph, n, camera22, blob, p, density, fb_gamma, fb_linear, ctr_out, build_stats, nullptr);
}

// Same renderer over traversal arrays produced elsewhere (the GPU build, read back with rt_octree_debug_read).
int hs_render_with_tree(const hs_sphere *sph, int n, const float *camera22, const hs_params *p, float *fb_gamma,
                        hs_counters *ctr_out, const void *nodes, int node_count, const void *node_ext, const void *cells,
                        int cell_count, const void *cell_ext, const void *vox_start, const void *vox_refs, const void *big_refs) {
    TreeView tv;
    memset(&tv, 0, sizeof tv);
    tv.nodes = static_cast<const TreeNode *>(nodes); tv.node_ext = static_cast<const TreeExtent *>(node_ext);
    tv.cells = static_cast<const CellGrid *>(cells); tv.cell_ext = static_cast<const TreeExtent *>(cell_ext);
    tv.vox_start = static_cast<const uint32_t *>(vox_start); tv.vox_refs = static_cast<const uint32_t *>(vox_refs);
    tv.big_refs = static_cast<const uint32_t *>(big_refs);
    tv.node_count = node_count; tv.cell_count = cell_count;
    return render_core(sph, n, camera22, nullptr, p, 0.f, fb_gamma, nullptr, ctr_out, nullptr, &tv);
}

// Host-built traversal arrays, for comparison with the ones the GPU build produces (rt_octree_debug_read):
// which = 0 nodes, 1 node_ext, 2 cells, 3 cell_ext, 4 vox_start, 5 vox_refs, 6 big_refs.  Returns bytes.
size_t hs_tree_dump(const hs_sphere *sph, int n, const void *blob, int spl, float density, int which, void *out, size_t cap) {
    std::vector<float4> geom((size_t)n);
    std::vector<int> tag((size_t)n);
    for (int i = 0; i < n; i++) { geom[(size_t)i] = make_float4(sph[i].cx, sph[i].cy, sph[i].cz, sph[i].radius); tag[(size_t)i] = sph[i].mat; }
    HostTree T;
    build_host_tree(geom, tag, static_cast<const int32_t *>(blob), spl, density, T);
    const void *src = nullptr;
    size_t bytes = 0;
    switch (which) {
        case 0: src = T.nodes.data(); bytes = T.nodes.size() * sizeof(TreeNode); break;
        case 1: src = T.node_ext.data(); bytes = T.node_ext.size() * sizeof(TreeExtent); break;
        case 2: src = T.cells.data(); bytes = T.cells.size() * sizeof(CellGrid); break;
        case 3: src = T.cell_ext.data(); bytes = T.cell_ext.size() * sizeof(TreeExtent); break;
        case 4: src = T.vox_start.data(); bytes = T.vox_start.size() * 4; break;
        case 5: src = T.vox_refs.data(); bytes = T.vox_refs.size() * 4; break;
        case 6: src = T.big_refs.data(); bytes = T.big_refs.size() * 4; break;
    }
    if (out && cap >= bytes && bytes) memcpy(out, src, bytes);
    return bytes;
}

}  // extern "C"
