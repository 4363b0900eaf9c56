"""2-rank NCCL worker for tests/test_configs_gpu.py::test_two_rank_nccl_sharding (one process per GPU).
Every rank also renders the unsharded result on its own GPU and compares bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    pkg = entry.import_package()
    from cloud_renderer_b200 import scene as sc, sharding as sh
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)

    # a 512^3 volume (the smallest that Z-slab sharding is offered for) with a small image and fewer billboards
    s = sc.make_scene("C4", boards=3000, size=(1920, 1080), frame=2)
    s.tp.sampler = pkg.SAMPLER_TEXTURE
    D, L = s.vol.dimension, s.vol.levels

    ref = pkg.Renderer(local)                                  # unsharded, on this GPU
    ref.set_scene(s); ref.voxelize()
    ref_img = ref.cone_trace(fmt=pkg.IMAGE_RGBA8).copy()
    ref_chain = ref.read_chain()
    ref.close()

    r = pkg.Renderer(local, stream.cuda_stream)
    r.set_scene(s)
    ex = sh.SlabExchange(torch, dist, r, D, L, rank, world, dev)
    r.set_tile_row_interleave(rank, world)
    for frame in range(2):                                     # twice: steady state reuses every buffer
        r.set_billboards(s.board_pos, s.board_scale)
        r.voxelize()
        ex.exchange()
        part = np.full_like(ref_img, 7)
        r.cone_trace(part, pkg.IMAGE_RGBA8)
    chain = r.read_chain()
    assert np.array_equal(chain, ref_chain), f"rank {rank}: chain after the slab exchange differs from the unsharded chain"
    mine = np.zeros(s.height, dtype=bool)
    for a, b in sh.tile_rows_of_rank(s.height, rank, world):
        mine[a:b] = True
    assert (part[~mine] == 7).all(), "rows of another rank were written"
    assert np.array_equal(part[mine], ref_img[mine]), f"rank {rank}: interleaved rows differ from the unsharded image"
    # assemble the image over NCCL and compare on every rank
    t = torch.from_numpy(np.where(mine[:, None, None], part, 0).astype(np.uint8)).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)                   # disjoint rows: the sum is the assembled image
    assert np.array_equal(t.cpu().numpy(), ref_img), "assembled image differs"
    r.close()

    # frames round-robin (C3/C5 sharding): frame g on rank g % world equals the same frame rendered alone
    r = pkg.Renderer(local, stream.cuda_stream)
    frames = [sc.make_scene("C2", frame=g, view=g) for g in range(rank, 4, world)]
    for f in frames:
        f.tp.sampler = pkg.SAMPLER_TEXTURE
        r.set_scene(f); r.voxelize()
        img = torch.from_numpy(r.cone_trace(fmt=pkg.IMAGE_RGBA8).copy()).to(dev)
        imgs = [torch.empty_like(img) for _ in range(world)]
        dist.all_gather(imgs, img)                             # rank q's frame is g = k*world + q
        g0 = f.meta["frame"] - rank
        for q in range(world):
            fq = sc.make_scene("C2", frame=g0 + q, view=g0 + q)
            fq.tp.sampler = pkg.SAMPLER_TEXTURE
            r.set_scene(fq); r.voxelize()
            assert np.array_equal(r.cone_trace(fmt=pkg.IMAGE_RGBA8), imgs[q].cpu().numpy()), f"frame {g0 + q} differs between ranks"
    r.close()
    dist.barrier()
    if rank == 0:
        print("NCCL_SHARDING_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
