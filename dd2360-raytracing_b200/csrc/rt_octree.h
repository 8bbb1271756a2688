// rt_octree.h — host-side handle of the GPU octree build (rt_octree.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "rt_types.h"

namespace rt {

struct BuildPlanes {
    float p[3][kPlanes];
};
struct BuildCounts {
    int node_count, cell_count;
    uint32_t total_voxels;
};

class OctreeBuilder {
public:
    OctreeBuilder();
    ~OctreeBuilder();
    OctreeBuilder(const OctreeBuilder &) = delete;
    OctreeBuilder &operator=(const OctreeBuilder &) = delete;

    // replaces D2H(spheres) + buildOctree + H2D(Octree), main.cu:405-415
    cudaError_t build(cudaStream_t st, const float4 *geom, const int *tag, int n, int spl, float density);
    static size_t reference_bytes(int spl);
    // the tree in the reference's own layout (acceleration_structure.h:23-61), assembled on the GPU on demand
    cudaError_t export_reference(cudaStream_t st, void *host_blob, size_t bytes);
    TreeView view() const;
    // test hook: copy one internal array to the host (0 nodes, 1 node_ext, 2 cells, 3 cell_ext, 4 vox_start,
    // 5 vox_refs, 6 big_refs); returns the byte size when host == nullptr
    size_t debug_read(cudaStream_t st, int which, void *host, size_t cap) const;

    bool built = false, blob_valid = false;
    int spl = 0, n_spheres = 0, leaf_count_h = 0;
    uint32_t E = 0, total_refs = 0;
    BuildCounts counts{};
    unsigned long long stats_h[4] = {0, 0, 0, 0};   // stored entries, dropped_full, dropped_outside, -
    BuildPlanes planes;

private:
    struct {
        uint32_t *ranges, *ent_count, *ent_off, *keys, *vals, *keys_sorted, *vals_sorted, *cell_count, *cell_start;
        uint8_t *entry_flag;
        CellGrid *raw;
        uint32_t *big_raw, *nvox;
        unsigned long long *stats;
        TreeNode *nodes;
        TreeExtent *node_ext;
        CellGrid *cells;
        TreeExtent *cell_ext;
        uint32_t *big_refs;
        int *node_of_potential, *dense_of_morton;
        BuildCounts *counts;
        uint32_t *vox_count, *vox_start, *vox_refs;
        uint8_t *cub_tmp;
        int *leaf_index;
        uint8_t *blob;
    } d;
    struct {
        size_t ranges, ent_count, ent_off, keys, vals, keys_sorted, vals_sorted, entry_flag, vox_count, vox_start,
            vox_refs, cub, blob;
    } cap;
};

}  // namespace rt
