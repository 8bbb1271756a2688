"""Frame sharding over ranks (one process per GPU, torch.distributed for the plumbing).

The reference is single-GPU.  A frame partitions naturally:
  * tiles : rank g renders the interleaved 8x4 pixel tiles {g, g+G, g+2G, ...} with ALL ns samples and the unmodified
            per-pixel stream, zeros elsewhere.  The sum over ranks is bit-identical to the 1-GPU frame.
  * spp   : rank g renders ns/G samples of every pixel from streams seeded 1984 + pixel_index + g*num_pixels.
            Statistically equivalent, not bit-identical (per-pixel samples are chained through the RNG, D8).
Either way the partial LINEAR radiance buffers are combined with ONE collective (reduce-sum to rank 0 over
NCCL/NVLink), then rank 0 applies /ns and sqrt (main.cu:111-114).  There is no other exchange on this path.
"""
from __future__ import annotations

import numpy as np

from . import SHARD_NONE, SHARD_SPP, SHARD_TILES, RayTracer

TILE_W, TILE_H = 8, 4


def tile_owner_map(nx: int, ny: int, world: int) -> np.ndarray:
    """[ny, nx] int32: which rank owns each pixel under RT_SHARD_TILES (tile index modulo world size)."""
    tx = (nx + TILE_W - 1) // TILE_W
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    tile = (jj // TILE_H) * tx + (ii // TILE_W)
    return (tile % world).astype(np.int32)


def spp_share(ns: int, rank: int, world: int) -> int:
    """Samples per pixel rank `rank` traces under RT_SHARD_SPP."""
    return ns // world + (1 if rank < ns % world else 0)


def reduce_frame(accum, dist, dst: int = 0):
    """The one collective of the path: sum the ranks' linear radiance buffers onto `dst` (in place)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum


def render_sharded(rt: RayTracer, accum, fb, nx: int, ny: int, ns: int, use_octree: bool, rank: int, world: int,
                   mode: int = SHARD_TILES, dist=None, want_stats: bool = True):
    """Render this rank's shard into `accum` (torch CUDA float32 tensor [ny, nx, 3]), reduce to rank 0, finalise into
    `fb` on rank 0.  Returns the per-rank render stats."""
    args = rt.args(nx, ny, ns, use_octree, shard_mode=mode if world > 1 else SHARD_NONE, shard_rank=rank, shard_count=world)
    st = rt.render_accumulate(args, accum.data_ptr(), want_stats=want_stats)
    reduce_frame(accum, dist)
    if rank == 0:
        rt.finalize(accum.data_ptr(), fb.data_ptr(), nx, ny, ns)
    return st


__all__ = ["tile_owner_map", "spp_share", "reduce_frame", "render_sharded", "SHARD_TILES", "SHARD_SPP", "SHARD_NONE"]
