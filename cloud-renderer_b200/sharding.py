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


class PackedSlabs:
    """ONE collective for several slab-partitioned buffers.  Every buffer in `parts` (flat uint8 tensors: the occupancy
    bits, chain levels 1..4) is split into `world` equal contiguous Z-slabs; rank r packs its slab of each into one chunk,
    a single in-place all_gather_into_tensor moves all chunks, and one strided copy per buffer puts the slabs back."""

    def __init__(self, torch, parts, rank, world):
        self.parts, self.rank, self.world = parts, rank, world
        self.sizes = [t.numel() // world for t in parts]
        for t, sz in zip(parts, self.sizes):
            assert sz * world == t.numel()
        self.chunk = sum(self.sizes)
        self.packed = torch.empty(world * self.chunk, dtype=torch.uint8, device=parts[0].device)
        self.mine = self.packed[rank * self.chunk:(rank + 1) * self.chunk]

    def pack(self):
        off = 0
        for t, sz in zip(self.parts, self.sizes):
            self.mine[off:off + sz].copy_(t[self.rank * sz:(self.rank + 1) * sz])
            off += sz

    def unpack(self):
        pv = self.packed.view(self.world, self.chunk)
        off = 0
        for t, sz in zip(self.parts, self.sizes):
            t.view(self.world, sz).copy_(pv[:, off:off + sz])
            off += sz

    def exchange(self, dist, group=None):
        self.pack()
        dist.all_gather_into_tensor(self.packed, self.mine, group=group)      # the ONE collective
        self.unpack()

    def bytes_received(self):
        return (self.world - 1) * self.chunk


class SlabExchange:
    """C4 (D >= 512): Z-slab voxelize + slab-local mips per rank, then ONE all-gather of the occupancy bits and chain
    levels 1..4 (level 0 is NOT shipped: 8x the bits, every rank re-expands it from them), replicated top levels."""

    def __init__(self, torch, dist, renderer, D, L, rank, world, device):
        self.dist, self.r, self.L = dist, renderer, L
        z0, z1 = z_slab(D, rank, world)
        renderer.set_z_slab(z0, z1)
        renderer.voxelize()                                   # allocates bits + chain
        renderer.sync()
        self.nloc = slab_local_levels(L)
        p, n = renderer.volume_bits_ptr()
        parts = [torch.as_tensor(DeviceBytes(p, n), device=device)]
        for l in range(1, self.nloc):
            p, n = renderer.volume_level_ptr(l)
            parts.append(torch.as_tensor(DeviceBytes(p, n), device=device))
        self.packed = PackedSlabs(torch, parts, rank, world)

    def exchange(self):
        self.packed.exchange(self.dist)
        if self.L > self.nloc:
            self.r.finish_mips(self.nloc)
        else:
            self.r.finish_mips(self.L)                        # nothing to reduce: marks the chain as changed

    def describe(self):
        return {"collectives_per_frame": 1, "bytes_received_per_rank": self.packed.bytes_received(),
                "shipped": "occupancy bits + chain levels 1..%d" % (self.nloc - 1)}


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
