"""N>1 host logic on CPU: world_size 2 over gloo + pure partition properties."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partitions(pkg):
    from cloud_renderer_b200 import sharding as sh
    for H in (96, 720, 1080, 2160, 4320, 17):
        for world in (1, 2, 4, 8):
            bands = [sh.row_range(H, r, world) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == H
            assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
            assert all(a % 16 == 0 or a == H for a, _ in bands)
            sizes = [b - a for a, b in bands]
            assert max(sizes) - min(sizes) < 32 or H < 16 * world      # one tile row + the partial last tile
    for D in (128, 256, 512):
        for world in (1, 2, 4, 8):
            slabs = [sh.z_slab(D, r, world) for r in range(world)]
            assert slabs[0][0] == 0 and slabs[-1][1] == D and all(a % 16 == 0 for a, _ in slabs)
    with pytest.raises(ValueError):
        sh.z_slab(64, 0, 8)                     # 8 slices per rank: thinner than a mip brick
    assert sh.frames_of_rank(3, 1, 4) == [1, 5, 9]
    assert sh.slab_local_levels(9) == 5 and sh.slab_local_levels(4) == 4


def test_world_size_2_gloo():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "_sharding_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and "SHARDING_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
