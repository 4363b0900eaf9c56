"""world_size-2 gloo worker for tests/test_sharding_cpu.py: the N>1 host logic on CPU.
The oracle stands in for the per-rank compute (this is a test; the product never does that)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    entry.import_package()
    from cloud_renderer_b200 import scene as sc, sharding as sh
    orc = entry.import_oracle()

    s = sc.make_scene("small")                         # D = 64, L = 5
    D, L = s.vol.dimension, s.vol.levels
    _, _, l0 = orc.voxelize(s, want_posmap=False)
    full = orc.mips(l0, L)

    # --- Z-slab scheme: own slab of the slab-local levels, one in-place all-gather, replicated top levels
    z0, z1 = sh.z_slab(D, rank, world)
    nloc = sh.slab_local_levels(L)
    mine0 = np.zeros_like(l0)
    mine0[z0:z1] = l0[z0:z1]                           # what crn_voxelize with crn_set_z_slab produces
    levels = []
    for l in range(nloc):
        size = D >> l
        t = torch.zeros(size ** 3, dtype=torch.uint8)
        lo, hi = z0 >> l, z1 >> l
        src = orc.chain_level(orc.mips(mine0, L), D, l)
        t.view(size, size, size)[lo:hi] = torch.from_numpy(src[lo:hi].copy())
        levels.append(t)
    views = [sh.slab_view(t, rank, world) for t in levels]
    sh.all_gather_levels(dist, views)
    for l in range(nloc):
        assert np.array_equal(levels[l].numpy(), orc.chain_level(full, D, l).ravel()), f"level {l} differs after the gather"
    top = orc.mips(levels[nloc - 1].numpy().reshape((D >> (nloc - 1),) * 3), L - nloc + 1)     # crn_finish_mips
    off = sum((D >> l) ** 3 for l in range(nloc - 1))
    assert np.array_equal(top, full[off:]), "replicated top levels differ"

    # --- the same exchange as ONE collective (PackedSlabs): bits + levels 1..nloc-1 packed into one chunk per rank
    bits_full = np.packbits((l0 > 0).ravel(), bitorder="little")
    bits_mine = np.zeros_like(bits_full)
    per = bits_full.size // world
    bits_mine[rank * per:(rank + 1) * per] = bits_full[rank * per:(rank + 1) * per]
    parts = [torch.from_numpy(bits_mine.copy())]
    for l in range(1, nloc):
        size = D >> l
        t = torch.zeros(size ** 3, dtype=torch.uint8)
        lo, hi = z0 >> l, z1 >> l
        src = orc.chain_level(orc.mips(mine0, L), D, l)
        t.view(size, size, size)[lo:hi] = torch.from_numpy(src[lo:hi].copy())
        parts.append(t)
    ps = sh.PackedSlabs(torch, parts, rank, world)
    ps.exchange(dist)
    assert np.array_equal(parts[0].numpy(), bits_full), "bits differ after the packed gather"
    for l in range(1, nloc):
        assert np.array_equal(parts[l].numpy(), orc.chain_level(full, D, l).ravel()), f"packed gather: level {l} differs"
    assert ps.bytes_received() == (world - 1) * sum(p.numel() // world for p in parts)

    # --- row bands of the trace: disjoint cover, and the assembled image equals the unsharded one
    H = s.height
    r0, r1 = sh.row_range(H, rank, world)
    img, _, _ = orc.cone_trace(s, full, rows=(r0, r1), want_u8=False)
    bands = [None] * world
    dist.all_gather_object(bands, (r0, r1, img[r0:r1].copy()))
    whole, _, _ = orc.cone_trace(s, full, want_u8=False)
    cover = np.zeros(H, dtype=np.int32)
    asm = np.zeros_like(whole)
    for a, b, part in bands:
        cover[a:b] += 1
        asm[a:b] = part
    assert (cover == 1).all() and np.array_equal(asm, whole)

    # --- frames round-robin: every global frame exactly once
    mine = sh.frames_of_rank(5, rank, world)
    allf = [None] * world
    dist.all_gather_object(allf, mine)
    assert sorted(sum(allf, [])) == list(range(5 * world))
    dist.barrier()
    if rank == 0:
        print("SHARDING_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
