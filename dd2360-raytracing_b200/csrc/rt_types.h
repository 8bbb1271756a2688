// rt_types.h — plain-data layouts shared by the host code, the build kernels and the render kernel.
#pragma once
#include <stdint.h>
#include <vector_types.h>

namespace rt {

// ---- reference Octree layout constants (acceleration_structure.h:11-15) ----
constexpr int kTreeHeight = 3;
constexpr int kNumberNodes = 585;         // 1 + 8 + 64 + 512
constexpr int kNumberLeafs = 4096;
constexpr int kNodeInts = 15;             // OctNode: level, AABB (6 floats), children[8]
constexpr int kCells = 512;               // level-3 cells of the fixed 8x8x8 subdivision
constexpr int kPlanes = 9;                // slab planes per axis at level 3

// Camera as the render kernel consumes it (camera.h:51-58): 22 floats, kept in constant memory.
struct CameraData {
    float origin[3], lower_left_corner[3], horizontal[3], vertical[3], u[3], v[3], w[3];
    float lens_radius;
};

// ---- traversal structure (internal; NOT the reference layout) ------------------------------------------------
// 32-byte packed octree node.  The reference AABB of a node is implied by (level, ix, iy, iz) on the fixed grid,
// so the node only carries its children and the bounding box of what it actually contains.
//   child[k]  : node index of child k (k = (x_high<<2)|(y_high<<1)|z_high, acceleration_structure.h:150-165);
//               for a level-3 node: k = 0 holds the cell-descriptor index; 0xFFFF = absent
//   ext_*     : AABB of the spheres stored below this node, half-precision-free: uint16 fixed point over the
//               padded root box would lose exactness we do not need here, so they live in TreeExtent instead.
struct alignas(32) TreeNode {
    uint16_t child[8];      // 16 B
    uint8_t level, ix, iy, iz;   // integer coordinates at this node's own level (0..(1<<level)-1)
    uint32_t first_cell;    // level-3: index into cells[]; else unused
    uint32_t pad[2];
};
static_assert(sizeof(TreeNode) == 32, "TreeNode must stay 32 bytes");

// conservative AABB of the content of a node / cell (padded at build time); 24 bytes + pad = 32
struct alignas(16) TreeExtent {
    float lo[3], hi[3];
    float pad[2];
};

// Per level-3 cell: a uniform sub-grid over the bounding box of the cell's stored SMALL spheres, plus a short
// list of BIG spheres (radius > kBigRadiusFrac * longest cell edge) that are tested directly when the cell is
// visited, so that one huge sphere does not coarsen the grid of a thousand small ones.
constexpr int kMaxBigPerCell = 16;
constexpr float kBigRadiusFrac = 0.25f;
struct alignas(16) CellGrid {
    float org[3];           // grid origin (min corner)
    uint32_t vox_base;      // first voxel of this grid in vox_start[]
    float inv_vs[3];        // 1 / voxel size
    uint32_t dims;          // nx | ny<<10 | nz<<20   (each <= 1023); 0 = no grid (big spheres only)
    float vs[3];            // voxel size
    uint32_t big;           // big_begin << 8 | big_count
    float hi[3];            // grid max corner
    uint32_t morton;        // which level-3 cell this is
};
static_assert(sizeof(CellGrid) == 64, "CellGrid is 64 bytes");

struct SceneView {
    const float4 *geom;     // {cx, cy, cz, radius} per sphere
    const float4 *matl;     // {albedo.xyz, param}
    const int *tag;         // RT_MAT_*
    int n;
};

struct TreeView {
    const TreeNode *nodes;      // node_count entries, node 0 = root
    const TreeExtent *node_ext; // per node
    const CellGrid *cells;      // per existing level-3 cell
    const TreeExtent *cell_ext;
    const uint32_t *vox_start;  // total_voxels + 1
    const uint32_t *vox_refs;   // sphere indices
    const uint32_t *big_refs;   // kMaxBigPerCell slots per cell
    int node_count, cell_count;
    float planes[3][kPlanes];   // slab plane coordinates per axis (exact floats of the reference subdivision)
};

}  // namespace rt
