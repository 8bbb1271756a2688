// rt_octree.cu — parallel GPU build of (a) the traversal structure the render kernel uses and (b) the tree in the
// reference's own memory layout, bit-exact against the serial buildOctree (acceleration_structure.h:195-217).
//
// Pipeline (one stream, no host round trip except two size read-backs needed to allocate):
//   k_classify     sphere -> per-axis level-3 slab ranges (reference `intersects` semantics), entry count,
//                  per-cell histogram (shared-memory privatised)
//   scan           entry offsets                                             (cub::DeviceScan)
//   k_emit         (Morton cell key, sphere index) pairs in sphere order
//   radix sort     stable, 9 key bits -> cell-major, sphere-ascending         (cub::DeviceRadixSort)
//   k_cell_grid    per cell: big/small split, bounding box of the stored list, sub-grid dimensions
//   k_assemble     one block: first-touch numbering of nodes (reference order), packed 32-byte nodes,
//                  content extents bottom-up, voxel bases
//   k_vox_count / scan / k_vox_fill / k_vox_sort   sphere-surface x voxel incidence lists
//   (on export)    k_leaf_number, k_blob_nodes, k_blob_leaves -> reference `Octree` blob
// CUB (shipped inside the CUDA toolkit, header-only) provides the device-wide scan and radix sort; everything
// else is hand-written.
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <limits.h>
#include <stdio.h>

#include "rt_build.cuh"
#include "rt_octree.h"

namespace rt {

#define RT_CUDA(x)                          \
    do {                                    \
        cudaError_t e_ = (x);               \
        if (e_ != cudaSuccess) return e_;   \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
__global__ void k_classify(const float4 *__restrict__ geom, int n, BuildPlanes P, uint32_t *__restrict__ ranges,
                           uint32_t *__restrict__ ent_count, uint32_t *__restrict__ cell_count,
                           unsigned long long *__restrict__ dropped_outside) {
    __shared__ uint32_t hist[kCells];
    for (int k = threadIdx.x; k < kCells; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        uint32_t packed = 0, cnt = 0;
        if (i >= 1) {   // ground sphere (index 0) is never inserted (acceleration_structure.h:208)
            const float4 s = geom[i];
            const AxisRange rx = axis_range(P.p[0], s.x, s.w, true);
            const AxisRange ry = axis_range(P.p[1], s.y, s.w, false);
            const AxisRange rz = axis_range(P.p[2], s.z, s.w, false);
            if (rx.lo <= rx.hi && ry.lo <= ry.hi && rz.lo <= rz.hi) {
                cnt = (uint32_t)((rx.hi - rx.lo + 1) * (ry.hi - ry.lo + 1) * (rz.hi - rz.lo + 1));
                packed = 1u << 31 | rx.lo | rx.hi << 4 | ry.lo << 8 | ry.hi << 12 | rz.lo << 16 | rz.hi << 20;
                for (int x = rx.lo; x <= rx.hi; x++)
                    for (int y = ry.lo; y <= ry.hi; y++)
                        for (int z = rz.lo; z <= rz.hi; z++) atomicAdd(&hist[morton_of(x, y, z)], 1u);
            } else {
                atomicAdd(dropped_outside, 1ull);   // ":105-108 Why is sphere not in range of nodes AABB?"
            }
        }
        ranges[i] = packed;
        ent_count[i] = cnt;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kCells; k += blockDim.x)
        if (hist[k]) atomicAdd(&cell_count[k], hist[k]);
}

__global__ void k_emit(const uint32_t *__restrict__ ranges, const uint32_t *__restrict__ ent_off, int n,
                       uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = ranges[i];
    if (!(p >> 31)) return;
    uint32_t o = ent_off[i];
    const int xl = p & 15, xh = (p >> 4) & 15, yl = (p >> 8) & 15, yh = (p >> 12) & 15, zl = (p >> 16) & 15, zh = (p >> 20) & 15;
    for (int x = xl; x <= xh; x++)
        for (int y = yl; y <= yh; y++)
            for (int z = zl; z <= zh; z++) {
                keys[o] = (uint32_t)morton_of(x, y, z);
                vals[o] = (uint32_t)i;
                o++;
            }
}

// exclusive scan of the 512 cell counts (one block of 512 threads)
__global__ void k_cell_scan(const uint32_t *__restrict__ cell_count, uint32_t *__restrict__ cell_start) {
    __shared__ uint32_t s[kCells];
    const int t = threadIdx.x;
    s[t] = cell_count[t];
    __syncthreads();
    for (int off = 1; off < kCells; off <<= 1) {
        uint32_t v = t >= off ? s[t - off] : 0;
        __syncthreads();
        s[t] += v;
        __syncthreads();
    }
    cell_start[t + 1] = s[t];
    if (t == 0) cell_start[0] = 0;
}

// One block per level-3 cell: split big/small, bound the stored list, choose the sub-grid.
__global__ void k_cell_grid(const float4 *__restrict__ geom, const int *__restrict__ tag,
                            const uint32_t *__restrict__ vals, const uint32_t *__restrict__ cell_start, int spl,
                            BuildPlanes P, float density, uint8_t *__restrict__ entry_flag, CellGrid *__restrict__ raw,
                            uint32_t *__restrict__ big_refs, uint32_t *__restrict__ nvox,
                            unsigned long long *__restrict__ stats /* [0]=stored entries, [1]=dropped_full */) {
    const int m = blockIdx.x;
    const uint32_t b = cell_start[m], e = cell_start[m + 1];
    const uint32_t count = e - b;
    const uint32_t cap = 8u * (uint32_t)spl;                       // 8 buckets of SPHERES_PER_LEAF (:110-136)
    const uint32_t stored = count < cap ? count : cap;
    __shared__ uint32_t s_big_n, s_live;
    __shared__ uint32_t s_big[kMaxBigPerCell];
    __shared__ float s_lo[3][32], s_hi[3][32];
    if (threadIdx.x == 0) { s_big_n = 0; s_live = 0; }
    __syncthreads();
    int ix, iy, iz;
    morton_to_xyz(m, ix, iy, iz);
    const float ex = P.p[0][ix + 1] - P.p[0][ix], ey = P.p[1][iy + 1] - P.p[1][iy], ez = P.p[2][iz + 1] - P.p[2][iz];
    const float big_r = kBigRadiusFrac * fmaxf(ex, fmaxf(ey, ez));
    // phase A: flags
    for (uint32_t k = threadIdx.x; k < count; k += blockDim.x) {
        uint8_t f = 2;                                              // dropped by the reference, or undefined slot
        if (k < stored) {
            const int idx = (int)vals[b + k];
            if (tag[idx] >= 0) {
                f = 0;
                if (geom[idx].w > big_r) {
                    const uint32_t slot = atomicAdd(&s_big_n, 1u);
                    if (slot < (uint32_t)kMaxBigPerCell) { s_big[slot] = (uint32_t)idx; f = 1; }
                }
            }
        }
        entry_flag[b + k] = f;
    }
    __syncthreads();
    // phase B: bounding box of the small stored spheres
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    uint32_t live = 0;
    for (uint32_t k = threadIdx.x; k < stored; k += blockDim.x) {
        if (entry_flag[b + k] != 0) continue;
        const float4 s = geom[vals[b + k]];
        const float r = s.w + sphere_pad(s.w);
        lo[0] = fminf(lo[0], s.x - r); hi[0] = fmaxf(hi[0], s.x + r);
        lo[1] = fminf(lo[1], s.y - r); hi[1] = fmaxf(hi[1], s.y + r);
        lo[2] = fminf(lo[2], s.z - r); hi[2] = fmaxf(hi[2], s.z + r);
        live++;
    }
    for (int o = 16; o > 0; o >>= 1) {
        for (int k = 0; k < 3; k++) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        live += __shfl_xor_sync(0xffffffffu, live, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        for (int k = 0; k < 3; k++) { s_lo[k][w] = lo[k]; s_hi[k][w] = hi[k]; }
        atomicAdd(&s_live, live);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int k = 0; k < 3; k++)
            for (int j = 1; j < nw; j++) { s_lo[k][0] = fminf(s_lo[k][0], s_lo[k][j]); s_hi[k][0] = fmaxf(s_hi[k][0], s_hi[k][j]); }
        CellGrid g;
        memset(&g, 0, sizeof g);
        g.morton = (uint32_t)m;
        uint32_t nb = s_big_n < (uint32_t)kMaxBigPerCell ? s_big_n : (uint32_t)kMaxBigPerCell;
        // ascending order keeps the traversal deterministic
        for (uint32_t a = 1; a < nb; a++) {
            uint32_t v = s_big[a];
            int j = (int)a - 1;
            while (j >= 0 && s_big[j] > v) { s_big[j + 1] = s_big[j]; j--; }
            s_big[j + 1] = v;
        }
        for (uint32_t a = 0; a < nb; a++) big_refs[m * kMaxBigPerCell + a] = s_big[a];
        g.big = nb;   // begin is filled in by k_assemble (dense cell index)
        float glo[3], ghi[3];
        for (int k = 0; k < 3; k++) { glo[k] = s_lo[k][0]; ghi[k] = s_hi[k][0]; }
        const uint32_t voxels = choose_grid(glo, ghi, s_live, density, g);
        raw[m] = g;
        nvox[m] = voxels;
        atomicAdd(&stats[0], (unsigned long long)stored);
        atomicAdd(&stats[1], (unsigned long long)(count - stored));
    }
}

// One block (1024 threads): number the nodes in the reference's creation order and emit the traversal arrays.
__global__ void k_assemble(const uint32_t *__restrict__ vals, const uint32_t *__restrict__ cell_start,
                           const float4 *__restrict__ geom, const CellGrid *__restrict__ raw,
                           const uint32_t *__restrict__ nvox, const uint32_t *__restrict__ big_refs_in, BuildPlanes P,
                           TreeNode *__restrict__ nodes, TreeExtent *__restrict__ node_ext, CellGrid *__restrict__ cells,
                           TreeExtent *__restrict__ cell_ext, uint32_t *__restrict__ big_refs_out,
                           int *__restrict__ node_of_potential, int *__restrict__ dense_of_morton,
                           BuildCounts *__restrict__ out) {
    __shared__ int first[kNumberNodes];      // first sphere index touching the potential node
    __shared__ int nidx[kNumberNodes];       // reference node index or -1
    __shared__ int dense[kCells];
    __shared__ uint32_t vbase[kCells];
    __shared__ float elo[kNumberNodes][3], ehi[kNumberNodes][3];
    const int t = threadIdx.x;
    // level 3
    if (t < kCells) {
        const uint32_t b = cell_start[t], e = cell_start[t + 1];
        first[level_base(3) + t] = e > b ? (int)vals[b] : INT_MAX;   // lists ascend, so the head is the first toucher
    }
    __syncthreads();
    if (t < 64) { int f = INT_MAX; for (int c = 0; c < 8; c++) f = min(f, first[level_base(3) + t * 8 + c]); first[level_base(2) + t] = f; }
    __syncthreads();
    if (t < 8) { int f = INT_MAX; for (int c = 0; c < 8; c++) f = min(f, first[level_base(2) + t * 8 + c]); first[level_base(1) + t] = f; }
    __syncthreads();
    if (t == 0) first[0] = -1;   // the root always exists and is created before any insertion (:200-206)
    __syncthreads();
    // creation order = sort by (first toucher, pre-order position); rank by counting (585^2 compares)
    int my_level = 0, my_path = 0;
    if (t < kNumberNodes) {
        my_level = t >= level_base(3) ? 3 : (t >= level_base(2) ? 2 : (t >= level_base(1) ? 1 : 0));
        my_path = t - level_base(my_level);
        int rank = -1;
        if (first[t] != INT_MAX) {
            const long long mykey = ((long long)first[t] << 12) | preorder_key(my_level, my_path);
            rank = 0;
            for (int j = 0; j < kNumberNodes; j++) {
                if (first[j] == INT_MAX) continue;
                const int lv = j >= level_base(3) ? 3 : (j >= level_base(2) ? 2 : (j >= level_base(1) ? 1 : 0));
                const long long key = ((long long)first[j] << 12) | preorder_key(lv, j - level_base(lv));
                rank += key < mykey;
            }
        }
        nidx[t] = rank;
        node_of_potential[t] = rank;
    }
    // dense cell numbering (Morton order) over cells that have anything to trace, and voxel bases
    if (t == 0) {
        int c = 0;
        uint32_t vb = 0;
        for (int m = 0; m < kCells; m++) {
            const bool live = nvox[m] > 0 || (raw[m].big & 0xff) > 0;
            dense[m] = live ? c++ : -1;
            vbase[m] = vb;
            vb += nvox[m];
        }
        out->cell_count = c;
        out->total_voxels = vb;
        int nc = 0;
        for (int j = 0; j < kNumberNodes; j++) nc += first[j] != INT_MAX;
        out->node_count = nc;
    }
    __syncthreads();
    // cells + their extents
    if (t < kCells) {
        dense_of_morton[t] = dense[t];
        const int id = level_base(3) + t;
        for (int k = 0; k < 3; k++) { elo[id][k] = 3e38f; ehi[id][k] = -3e38f; }
        if (dense[t] >= 0) {
            CellGrid g = raw[t];
            g.vox_base = vbase[t];
            const uint32_t nb = g.big & 0xff;
            g.big = ((uint32_t)dense[t] * kMaxBigPerCell) << 8 | nb;
            if (g.dims) for (int k = 0; k < 3; k++) { elo[id][k] = g.org[k]; ehi[id][k] = g.hi[k]; }
            for (uint32_t a = 0; a < nb; a++) {
                const uint32_t idx = big_refs_in[t * kMaxBigPerCell + a];
                big_refs_out[dense[t] * kMaxBigPerCell + a] = idx;
                const float4 s = geom[idx];
                const float r = s.w + sphere_pad(s.w);
                elo[id][0] = fminf(elo[id][0], s.x - r); ehi[id][0] = fmaxf(ehi[id][0], s.x + r);
                elo[id][1] = fminf(elo[id][1], s.y - r); ehi[id][1] = fmaxf(ehi[id][1], s.y + r);
                elo[id][2] = fminf(elo[id][2], s.z - r); ehi[id][2] = fmaxf(ehi[id][2], s.z + r);
            }
            cells[dense[t]] = g;
            TreeExtent x;
            for (int k = 0; k < 3; k++) { x.lo[k] = elo[id][k]; x.hi[k] = ehi[id][k]; }
            x.pad[0] = x.pad[1] = 0.f;
            cell_ext[dense[t]] = x;
        }
    }
    __syncthreads();
    for (int lv = 2; lv >= 0; lv--) {    // extents bottom-up
        const int cnt = 1 << (3 * lv);
        if (t < cnt) {
            const int id = level_base(lv) + t;
            for (int k = 0; k < 3; k++) { elo[id][k] = 3e38f; ehi[id][k] = -3e38f; }
            for (int c = 0; c < 8; c++) {
                const int ch = level_base(lv + 1) + t * 8 + c;
                for (int k = 0; k < 3; k++) { elo[id][k] = fminf(elo[id][k], elo[ch][k]); ehi[id][k] = fmaxf(ehi[id][k], ehi[ch][k]); }
            }
        }
        __syncthreads();
    }
    if (t < kNumberNodes && nidx[t] >= 0) {
        TreeNode nd;
        memset(&nd, 0, sizeof nd);
        nd.level = (uint8_t)my_level;
        int ix = 0, iy = 0, iz = 0;
        for (int l = 0; l < my_level; l++) {
            const int c = (my_path >> (3 * (my_level - 1 - l))) & 7;
            ix = (ix << 1) | (c >> 2); iy = (iy << 1) | ((c >> 1) & 1); iz = (iz << 1) | (c & 1);
        }
        nd.ix = (uint8_t)ix; nd.iy = (uint8_t)iy; nd.iz = (uint8_t)iz;
        if (my_level < 3) {
            for (int c = 0; c < 8; c++) {
                const int ch = nidx[level_base(my_level + 1) + my_path * 8 + c];
                nd.child[c] = ch >= 0 ? (uint16_t)ch : (uint16_t)0;    // 0 = absent, as in the reference
            }
        } else {
            nd.first_cell = dense[my_path] >= 0 ? (uint32_t)dense[my_path] : 0xffffffffu;
        }
        nodes[nidx[t]] = nd;
        TreeExtent x;
        for (int k = 0; k < 3; k++) { x.lo[k] = elo[t][k]; x.hi[k] = ehi[t][k]; }
        x.pad[0] = x.pad[1] = 0.f;
        node_ext[nidx[t]] = x;
    }
}

// enumerate the voxels of `g` crossed by the (padded) surface of sphere s
template <typename F>
__device__ __forceinline__ void for_each_voxel(const CellGrid &g, const float4 s, F f) {
    const int nx = (int)(g.dims & 1023u), ny = (int)((g.dims >> 10) & 1023u);
    const float pad = sphere_pad(s.w);
    int v0[3], v1[3];
    voxel_range(g, s, pad, v0, v1);
    for (int z = v0[2]; z <= v1[2]; z++)
        for (int y = v0[1]; y <= v1[1]; y++)
            for (int x = v0[0]; x <= v1[0]; x++) {
                float lo[3], hi[3];
                voxel_box(g, x, y, z, lo, hi);
                if (shell_hits_box(s, pad, lo, hi)) f(g.vox_base + (uint32_t)((z * ny + y) * nx + x));
            }
}

template <bool FILL>
__global__ void k_vox_pass(const float4 *__restrict__ geom, const uint32_t *__restrict__ keys,
                           const uint32_t *__restrict__ vals, const uint8_t *__restrict__ entry_flag, uint32_t E,
                           const int *__restrict__ dense_of_morton, const CellGrid *__restrict__ cells,
                           uint32_t *__restrict__ counter, const uint32_t *__restrict__ vox_start,
                           uint32_t *__restrict__ refs) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= E) return;
    if (entry_flag[p] != 0) return;
    const int dc = dense_of_morton[keys[p]];
    if (dc < 0) return;
    const CellGrid g = cells[dc];
    if (!g.dims) return;
    const uint32_t idx = vals[p];
    const float4 s = geom[idx];
    for_each_voxel(g, s, [&](uint32_t v) {
        const uint32_t slot = atomicAdd(&counter[v], 1u);
        if (FILL) refs[vox_start[v] + slot] = idx;
    });
}

__global__ void k_vox_sort(const uint32_t *__restrict__ vox_start, uint32_t total_voxels, uint32_t *__restrict__ refs) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= total_voxels) return;
    const uint32_t b = vox_start[v], e = vox_start[v + 1];
    for (uint32_t a = b + 1; a < e; a++) {
        const uint32_t x = refs[a];
        uint32_t j = a;
        while (j > b && refs[j - 1] > x) { refs[j] = refs[j - 1]; j--; }
        refs[j] = x;
    }
}

// ---- reference-layout blob ------------------------------------------------------------------------------------
// bucket (cell m, slot k) exists when k*spl < stored(m); its creator is the sphere with in-cell rank k*spl.
__global__ void k_leaf_number(const uint32_t *__restrict__ vals, const uint32_t *__restrict__ cell_start, int spl,
                              int *__restrict__ leaf_index /* [512*8] */, int *__restrict__ leaf_count) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;   // m*8 + k
    if (id >= kCells * 8) return;
    auto key_of = [&](int j) -> long long {
        const int m = j >> 3, k = j & 7;
        const uint32_t b = cell_start[m], e = cell_start[m + 1];
        const uint32_t count = e - b, cap = 8u * (uint32_t)spl;
        const uint32_t stored = count < cap ? count : cap;
        if ((uint32_t)k * (uint32_t)spl >= stored) return -1;
        return ((long long)vals[b + (uint32_t)k * (uint32_t)spl] << 12) | preorder_key(3, m);
    };
    const long long mykey = key_of(id);
    int rank = -1;
    if (mykey >= 0) {
        rank = 0;
        for (int j = 0; j < kCells * 8; j++) {
            const long long kj = key_of(j);
            rank += (kj >= 0 && kj < mykey);
        }
        atomicAdd(leaf_count, 1);
    }
    leaf_index[id] = rank >= 0 ? rank + 1 : 0;   // leaves[0] is never used (:58)
}

__global__ void k_blob_nodes(const int *__restrict__ node_of_potential, const int *__restrict__ leaf_index,
                             BuildPlanes P, int32_t *__restrict__ blob_nodes) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= kNumberNodes) return;
    const int ni = node_of_potential[t];
    if (ni < 0) return;
    const int lv = t >= level_base(3) ? 3 : (t >= level_base(2) ? 2 : (t >= level_base(1) ? 1 : 0));
    const int path = t - level_base(lv);
    int ix = 0, iy = 0, iz = 0;
    for (int l = 0; l < lv; l++) {
        const int c = (path >> (3 * (lv - 1 - l))) & 7;
        ix = (ix << 1) | (c >> 2); iy = (iy << 1) | ((c >> 1) & 1); iz = (iz << 1) | (c & 1);
    }
    const int sh = 3 - lv;
    int32_t *nd = blob_nodes + ni * kNodeInts;
    nd[0] = lv;
    float *bx = reinterpret_cast<float *>(nd + 1);
    bx[0] = P.p[0][ix << sh]; bx[1] = P.p[1][iy << sh]; bx[2] = P.p[2][iz << sh];
    bx[3] = P.p[0][(ix + 1) << sh]; bx[4] = P.p[1][(iy + 1) << sh]; bx[5] = P.p[2][(iz + 1) << sh];
    for (int c = 0; c < 8; c++) {
        int v;
        if (lv < 3) { const int ch = node_of_potential[level_base(lv + 1) + path * 8 + c]; v = ch >= 0 ? ch : 0; }
        else v = leaf_index[path * 8 + c];
        nd[7 + c] = v;
    }
}

__global__ void k_blob_leaves(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                              const uint32_t *__restrict__ cell_start, uint32_t E, int spl,
                              const int *__restrict__ leaf_index, int32_t *__restrict__ blob_leaves) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= E) return;
    const uint32_t m = keys[p];
    const uint32_t b = cell_start[m], e = cell_start[m + 1];
    const uint32_t pos = p - b, count = e - b, cap = 8u * (uint32_t)spl;
    if (pos >= cap) return;                                   // "Leaf nodes full" (:135)
    const uint32_t stored = count < cap ? count : cap;
    const uint32_t k = pos / (uint32_t)spl, slot = pos % (uint32_t)spl;
    int32_t *lf = blob_leaves + (size_t)leaf_index[m * 8 + k] * (size_t)(spl + 1);
    lf[slot] = (int32_t)vals[p];
    if (slot == 0) {
        const uint32_t left = stored - k * (uint32_t)spl;
        lf[spl] = (int32_t)(left < (uint32_t)spl ? left : (uint32_t)spl);
    }
}

// ---------------------------------------------------------------------------------------------------------------
static void make_planes(BuildPlanes &P) {
    // acceleration_structure.h:141-147,203: root (-11,0,-11)-(11,2,11), children by low + (high-low)/2 in float
    const float lo[3] = {-11.f, 0.f, -11.f}, hi[3] = {11.f, 2.f, 11.f};
    for (int a = 0; a < 3; a++) {
        float *p = P.p[a];
        p[0] = lo[a]; p[8] = hi[a];
        for (int step = 8; step > 1; step >>= 1)
            for (int i = 0; i < 8; i += step) {
                volatile float l = p[i], h = p[i + step];
                volatile float d = h - l;
                volatile float half = d / 2;
                p[i + step / 2] = l + half;
            }
    }
}

template <typename T>
static cudaError_t ensure(T *&ptr, size_t &cap, size_t need) {
    if (need <= cap) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    size_t want = need + need / 4 + 256;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&ptr), want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
}

OctreeBuilder::OctreeBuilder() { memset(&d, 0, sizeof d); memset(&cap, 0, sizeof cap); make_planes(planes); }

OctreeBuilder::~OctreeBuilder() {
    void *ptrs[] = {d.ranges, d.ent_count, d.ent_off, d.keys, d.vals, d.keys_sorted, d.vals_sorted, d.cell_count,
                    d.cell_start, d.entry_flag, d.raw, d.big_raw, d.nvox, d.stats, d.nodes, d.node_ext, d.cells,
                    d.cell_ext, d.big_refs, d.node_of_potential, d.dense_of_morton, d.counts, d.vox_count, d.vox_start,
                    d.vox_refs, d.cub_tmp, d.leaf_index, d.blob};
    for (void *p : ptrs) if (p) cudaFree(p);
}

cudaError_t OctreeBuilder::build(cudaStream_t st, const float4 *geom, const int *tag, int n, int spl_, float density) {
    spl = spl_;
    built = false;
    blob_valid = false;
    RT_CUDA(ensure(d.ranges, cap.ranges, (size_t)n + 1));
    RT_CUDA(ensure(d.ent_count, cap.ent_count, (size_t)n + 1));
    RT_CUDA(ensure(d.ent_off, cap.ent_off, (size_t)n + 1));
    if (!d.cell_count) {
        RT_CUDA(cudaMalloc(&d.cell_count, kCells * 4));
        RT_CUDA(cudaMalloc(&d.cell_start, (kCells + 1) * 4));
        RT_CUDA(cudaMalloc(&d.raw, kCells * sizeof(CellGrid)));
        RT_CUDA(cudaMalloc(&d.big_raw, kCells * kMaxBigPerCell * 4));
        RT_CUDA(cudaMalloc(&d.nvox, kCells * 4));
        RT_CUDA(cudaMalloc(&d.stats, 4 * 8));
        RT_CUDA(cudaMalloc(&d.nodes, kNumberNodes * sizeof(TreeNode)));
        RT_CUDA(cudaMalloc(&d.node_ext, kNumberNodes * sizeof(TreeExtent)));
        RT_CUDA(cudaMalloc(&d.cells, kCells * sizeof(CellGrid)));
        RT_CUDA(cudaMalloc(&d.cell_ext, kCells * sizeof(TreeExtent)));
        RT_CUDA(cudaMalloc(&d.big_refs, kCells * kMaxBigPerCell * 4));
        RT_CUDA(cudaMalloc(&d.node_of_potential, kNumberNodes * 4));
        RT_CUDA(cudaMalloc(&d.dense_of_morton, kCells * 4));
        RT_CUDA(cudaMalloc(&d.counts, sizeof(BuildCounts)));
        RT_CUDA(cudaMalloc(&d.leaf_index, kCells * 8 * 4 + 4));
    }
    RT_CUDA(cudaMemsetAsync(d.cell_count, 0, kCells * 4, st));
    RT_CUDA(cudaMemsetAsync(d.stats, 0, 4 * 8, st));
    RT_CUDA(cudaMemsetAsync(d.ent_count + n, 0, 4, st));
    const int tb = 256;
    k_classify<<<(n + tb - 1) / tb, tb, 0, st>>>(geom, n, planes, d.ranges, d.ent_count, d.cell_count, d.stats + 2);
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, d.ent_count, d.ent_off, n + 1, st);
    RT_CUDA(ensure(d.cub_tmp, cap.cub, tmp));
    cub::DeviceScan::ExclusiveSum(d.cub_tmp, tmp, d.ent_count, d.ent_off, n + 1, st);
    uint32_t E_h = 0;
    RT_CUDA(cudaMemcpyAsync(&E_h, d.ent_off + n, 4, cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    E = E_h;
    RT_CUDA(ensure(d.keys, cap.keys, (size_t)E + 1));
    RT_CUDA(ensure(d.vals, cap.vals, (size_t)E + 1));
    RT_CUDA(ensure(d.keys_sorted, cap.keys_sorted, (size_t)E + 1));
    RT_CUDA(ensure(d.vals_sorted, cap.vals_sorted, (size_t)E + 1));
    RT_CUDA(ensure(d.entry_flag, cap.entry_flag, (size_t)E + 1));
    k_emit<<<(n + tb - 1) / tb, tb, 0, st>>>(d.ranges, d.ent_off, n, d.keys, d.vals);
    if (E > 0) {
        tmp = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp, d.keys, d.keys_sorted, d.vals, d.vals_sorted, (int)E, 0, 9, st);
        RT_CUDA(ensure(d.cub_tmp, cap.cub, tmp));
        cub::DeviceRadixSort::SortPairs(d.cub_tmp, tmp, d.keys, d.keys_sorted, d.vals, d.vals_sorted, (int)E, 0, 9, st);
    }
    k_cell_scan<<<1, kCells, 0, st>>>(d.cell_count, d.cell_start);
    k_cell_grid<<<kCells, 256, 0, st>>>(geom, tag, d.vals_sorted, d.cell_start, spl, planes, density, d.entry_flag, d.raw,
                                       d.big_raw, d.nvox, d.stats);
    k_assemble<<<1, 1024, 0, st>>>(d.vals_sorted, d.cell_start, geom, d.raw, d.nvox, d.big_raw, planes, d.nodes, d.node_ext,
                                  d.cells, d.cell_ext, d.big_refs, d.node_of_potential, d.dense_of_morton, d.counts);
    RT_CUDA(cudaMemcpyAsync(&counts, d.counts, sizeof counts, cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaMemcpyAsync(stats_h, d.stats, 4 * 8, cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    const uint32_t V = counts.total_voxels;
    RT_CUDA(ensure(d.vox_count, cap.vox_count, (size_t)V + 2));
    RT_CUDA(ensure(d.vox_start, cap.vox_start, (size_t)V + 2));
    RT_CUDA(cudaMemsetAsync(d.vox_count, 0, ((size_t)V + 2) * 4, st));
    if (E > 0 && V > 0)
        k_vox_pass<false><<<(E + tb - 1) / tb, tb, 0, st>>>(geom, d.keys_sorted, d.vals_sorted, d.entry_flag, E,
                                                            d.dense_of_morton, d.cells, d.vox_count, nullptr, nullptr);
    tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, d.vox_count, d.vox_start, (int)V + 1, st);
    RT_CUDA(ensure(d.cub_tmp, cap.cub, tmp));
    cub::DeviceScan::ExclusiveSum(d.cub_tmp, tmp, d.vox_count, d.vox_start, (int)V + 1, st);
    uint32_t R_h = 0;
    RT_CUDA(cudaMemcpyAsync(&R_h, d.vox_start + V, 4, cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    total_refs = R_h;
    RT_CUDA(ensure(d.vox_refs, cap.vox_refs, (size_t)total_refs + 1));
    RT_CUDA(cudaMemsetAsync(d.vox_count, 0, ((size_t)V + 2) * 4, st));
    if (E > 0 && V > 0) {
        k_vox_pass<true><<<(E + tb - 1) / tb, tb, 0, st>>>(geom, d.keys_sorted, d.vals_sorted, d.entry_flag, E,
                                                           d.dense_of_morton, d.cells, d.vox_count, d.vox_start, d.vox_refs);
        k_vox_sort<<<(V + tb - 1) / tb, tb, 0, st>>>(d.vox_start, V, d.vox_refs);
    }
    RT_CUDA(cudaGetLastError());
    built = true;
    n_spheres = n;
    return cudaSuccess;
}

size_t OctreeBuilder::reference_bytes(int spl_) {
    return (size_t)kNumberNodes * kNodeInts * 4 + (size_t)(kNumberLeafs + 1) * (size_t)(spl_ + 1) * 4 + 8;
}

cudaError_t OctreeBuilder::export_reference(cudaStream_t st, void *host_blob, size_t bytes) {
    if (!built) return cudaErrorNotReady;
    const size_t need = reference_bytes(spl);
    if (bytes < need) return cudaErrorInvalidValue;
    RT_CUDA(ensure(d.blob, cap.blob, need));
    RT_CUDA(cudaMemsetAsync(d.blob, 0, need, st));          // `new Octree()` value-initialises
    int *leaf_count = d.leaf_index + kCells * 8;
    RT_CUDA(cudaMemsetAsync(leaf_count, 0, 4, st));
    k_leaf_number<<<(kCells * 8 + 255) / 256, 256, 0, st>>>(d.vals_sorted, d.cell_start, spl, d.leaf_index, leaf_count);
    int32_t *nodes = reinterpret_cast<int32_t *>(d.blob);
    int32_t *leaves = nodes + kNumberNodes * kNodeInts;
    k_blob_nodes<<<(kNumberNodes + 255) / 256, 256, 0, st>>>(d.node_of_potential, d.leaf_index, planes, nodes);
    if (E > 0)
        k_blob_leaves<<<(E + 255) / 256, 256, 0, st>>>(d.keys_sorted, d.vals_sorted, d.cell_start, E, spl, d.leaf_index, leaves);
    int lc = 0;
    RT_CUDA(cudaMemcpyAsync(&lc, leaf_count, 4, cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    leaf_count_h = lc + 1;                                   // leafCount starts at 1 (:60)
    const int32_t tail[2] = {counts.node_count, leaf_count_h};
    RT_CUDA(cudaMemcpyAsync(d.blob + need - 8, tail, 8, cudaMemcpyHostToDevice, st));
    RT_CUDA(cudaMemcpyAsync(host_blob, d.blob, need, cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    blob_valid = true;
    return cudaGetLastError();
}

size_t OctreeBuilder::debug_read(cudaStream_t st, int which, void *host, size_t cap_bytes) const {
    if (!built) return 0;
    const void *src = nullptr;
    size_t bytes = 0;
    switch (which) {
        case 0: src = d.nodes; bytes = (size_t)counts.node_count * sizeof(TreeNode); break;
        case 1: src = d.node_ext; bytes = (size_t)counts.node_count * sizeof(TreeExtent); break;
        case 2: src = d.cells; bytes = (size_t)counts.cell_count * sizeof(CellGrid); break;
        case 3: src = d.cell_ext; bytes = (size_t)counts.cell_count * sizeof(TreeExtent); break;
        case 4: src = d.vox_start; bytes = ((size_t)counts.total_voxels + 1) * 4; break;
        case 5: src = d.vox_refs; bytes = (size_t)total_refs * 4; break;
        case 6: src = d.big_refs; bytes = (size_t)counts.cell_count * kMaxBigPerCell * 4; break;
        default: return 0;
    }
    if (!host) return bytes;
    if (cap_bytes < bytes) return 0;
    if (bytes && cudaMemcpyAsync(host, src, bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) return 0;
    cudaStreamSynchronize(st);
    return bytes;
}

TreeView OctreeBuilder::view() const {
    TreeView v;
    memset(&v, 0, sizeof v);
    v.nodes = d.nodes; v.node_ext = d.node_ext; v.cells = d.cells; v.cell_ext = d.cell_ext;
    v.vox_start = d.vox_start; v.vox_refs = d.vox_refs; v.big_refs = d.big_refs;
    v.node_count = counts.node_count; v.cell_count = counts.cell_count;
    for (int a = 0; a < 3; a++) for (int i = 0; i < kPlanes; i++) v.planes[a][i] = planes.p[a][i];
    return v;
}

}  // namespace rt
