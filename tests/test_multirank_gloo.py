"""CPU tier: the N>1 host logic with world_size 2 over gloo — tile ownership, spp shares, and the one collective
(reduce-sum of linear radiance) reproducing the unsharded frame.  The per-shard 'renderer' here is the oracle
(this is a test of the sharding / reduction logic, not of the CUDA kernel)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %(root)r)
    import numpy as np, torch, torch.distributed as dist
    import __graft_entry__ as entry
    pkg = entry.load_package(); O = entry.load_oracle()
    from dd2360_raytracing_b200 import multigpu as mg
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
    rank, world = dist.get_rank(), dist.get_world_size()
    nx, ny, ns, n, spl = 52, 30, 3, 488, 30
    sph, _ = O.create_world(n); blob, _ = O.build_octree(sph, spl)
    cam = O.camera(nx, ny, O.ARITH_DEVICE)
    _, full, _ = O.render(sph, cam, O.make_params(nx, ny, ns, True, spl, O.ARITH_DEVICE), blob, want_linear=True)
    # tile shards: this rank keeps only the pixels it owns
    owner = mg.tile_owner_map(nx, ny, world)
    part = torch.from_numpy(np.where((owner == rank)[..., None], full, 0.0).astype(np.float32))
    mg.reduce_frame(part, dist)
    if rank == 0:
        assert np.array_equal(part.numpy().view(np.uint32), full.view(np.uint32)), "tile shards do not sum to the frame"
    # the path's one collective as bench.py runs it: reduce-scatter of the padded linear buffer, every rank finalises its own
    # slice and writes it into ONE host frame shared by the ranks (POSIX shared memory); 3 ranks' worth of padding is exercised
    # by a frame whose element count is odd
    total = nx * ny * 3
    S = mg.slice_elems(total, world)
    padded = torch.zeros(S * world, dtype=torch.float32)
    padded[:total] = torch.from_numpy(np.where((owner == rank)[..., None], full, 0.0).astype(np.float32)).reshape(-1)
    mine = torch.empty(S, dtype=torch.float32)
    mg.reduce_scatter_frame(padded, mine, dist)
    b, e = mg.slice_range(total, rank, world)
    assert (b, e) == (rank * S, min((rank + 1) * S, total))
    assert np.array_equal(mine.numpy()[: e - b].view(np.uint32), full.reshape(-1)[b:e].view(np.uint32)), "reduce-scatter slice differs"
    name = "rt_b200_test_%(port)d"
    if rank == 0:
        host = mg.SharedHostFrame(name, total * 4, True, torch)
    dist.barrier()
    if rank != 0:
        host = mg.SharedHostFrame(name, total * 4, False, torch)
    fin = np.sqrt(mine.numpy()[: e - b] * np.float32(1.0 / ns)).astype(np.float32)       # /ns and sqrt of the slice (main.cu:111-114)
    host.tensor[b:e].copy_(torch.from_numpy(fin))
    dist.barrier()
    if rank == 0:
        want = np.sqrt(full.reshape(-1) * np.float32(1.0 / ns)).astype(np.float32)
        assert np.array_equal(host.array.view(np.uint32), want.view(np.uint32)), "the shared host frame is not the finalised frame"
    dist.barrier()
    host.close()
    assert mg.slice_elems(7, 3) == 3 and mg.slice_range(7, 2, 3) == (6, 7) and mg.slice_range(7, 1, 3) == (3, 6)
    # spp shares add up and differ by at most one
    shares = [mg.spp_share(7, r, world) for r in range(world)]
    assert sum(shares) == 7 and max(shares) - min(shares) <= 1
    # every tile has exactly one owner, ownership is balanced
    counts = np.bincount(owner.ravel(), minlength=world)
    assert counts.sum() == nx * ny and abs(int(counts[0]) - int(counts[1])) <= 8 * 4 * ((nx + 7) // 8)
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_two_rank_gloo_tile_reduce(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o


def test_tile_owner_map_matches_kernel_convention():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    entry.load_package()
    from dd2360_raytracing_b200 import multigpu as mg
    m = mg.tile_owner_map(20, 9, 3)
    assert m.shape == (9, 20)
    # tiles are 8 wide, 4 high, numbered row-major; tile t belongs to rank t % world
    assert m[0, 0] == 0 and m[0, 8] == 1 and m[0, 16] == 2 and m[4, 0] == 0 and m[4, 8] == 1 and m[8, 16] == (2 * 3 + 2) % 3
