"""Host-side partitioning for one 8xB200 box.  One process per GPU.

Only what shards naturally is sharded:
  * frames / views  : round-robin over ranks, volume replicated, NO collective (C3, C5);
  * image rows      : contiguous bands of the cone trace, volume replicated, NO collective
                      for compute (C4's trace);
  * volume Z-slabs  : voxelize pass 2 + the slab-local mip levels per rank, then ONE
                      all-gather of the finished chain over NCCL/NVLink and a replicated
                      reduction of the few top levels (C4, D >= 512).
Every function here is pure index arithmetic or a torch.distributed call on buffers the
C-ABI exposes, so the same code runs under gloo on CPU in tests/test_sharding_cpu.py.
"""
from dataclasses import dataclass

SLAB_ALIGN = 16          # k_mips.cu bricks are 16 slices deep: levels 0..4 are slab-local
SLAB_LOCAL_LEVELS = 5


def frames_of_rank(n_frames_per_rank, rank, world):
    """global frame ids rendered by `rank`: k*world + rank (round-robin, weak scaling)"""
    return [k * world + rank for k in range(n_frames_per_rank)]


def row_range(height, rank, world, align=16):
    """contiguous band of image rows for `rank`; bands are multiples of the 16-pixel tile
    except the last, cover [0,height) exactly once and differ by at most one tile row."""
    tiles = (height + align - 1) // align
    base, extra = divmod(tiles, world)
    t0 = rank * base + min(rank, extra)
    t1 = t0 + base + (1 if rank < extra else 0)
    return min(t0 * align, height), min(t1 * align, height)


def tile_rows_of_rank(height, rank, world, tile=16):
    """load-balanced alternative to row_range: 16-pixel tile rows dealt round-robin (the cloud sits in the
    middle of the frame, so contiguous bands would leave the outer ranks idle).  -> list of (row0,row1)"""
    tiles = (height + tile - 1) // tile
    return [(t * tile, min(height, (t + 1) * tile)) for t in range(rank, tiles, world)]


def z_slab(dim, rank, world):
    """voxel slices [z0,z1) owned by `rank`; equal, SLAB_ALIGN-aligned slabs"""
    if dim % (world * SLAB_ALIGN):
        raise ValueError(f"dimension {dim} does not split into {world} slabs of a multiple of {SLAB_ALIGN} slices")
    t = dim // world
    return rank * t, (rank + 1) * t


def slab_local_levels(levels):
    return min(levels, SLAB_LOCAL_LEVELS)


@dataclass
class LevelView:
    """whole level `l` of the chain as a flat uint8 tensor + this rank's slab of it"""
    whole: object
    mine: object


def slab_view(flat_tensor, rank, world):
    """Z-slabs of an x-fastest volume (bytes or bits) are contiguous, equal chunks in rank order, so
    the flat level tensor IS the all-gather output and the rank's slab is the in-place input."""
    n = flat_tensor.numel()
    assert n % world == 0
    chunk = n // world
    return LevelView(flat_tensor, flat_tensor[rank * chunk:(rank + 1) * chunk])


def all_gather_levels(dist, views, group=None):
    """the ONE exchange of the Z-slab scheme: all-gather every slab-local level in place"""
    works = [dist.all_gather_into_tensor(v.whole, v.mine, group=group, async_op=True) for v in views]
    for w in works:
        w.wait()


class DeviceBytes:
    """zero-copy torch view of a device buffer the C-ABI owns (crn_volume_level_ptr)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def chain_tensors(torch, renderer, levels, device):
    """[bits, level 0, level 1, ...] as torch uint8 tensors aliasing the context's memory"""
    out = []
    p, n = renderer.volume_bits_ptr()
    out.append(torch.as_tensor(DeviceBytes(p, n), device=device))
    for l in range(levels):
        p, n = renderer.volume_level_ptr(l)
        out.append(torch.as_tensor(DeviceBytes(p, n), device=device))
    return out
