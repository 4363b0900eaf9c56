import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    p = entry.import_package()
    if not os.path.exists(p.LIB_PATH):
        entry.build()
    p.load_library()
    return p


@pytest.fixture(scope="session")
def scenes(pkg):
    from cloud_renderer_b200 import scene
    return scene


@pytest.fixture(scope="session")
def orc():
    o = entry.import_oracle()
    o.build()
    o.lib()
    return o


@pytest.fixture(scope="session")
def renderer(pkg):
    r = pkg.Renderer(0)
    yield r
    r.close()


def psnr(a, b, peak=1.0):
    mse = float(np.mean((np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)) ** 2))
    return 99.0 if mse == 0.0 else 10.0 * np.log10(peak * peak / mse)


def steady_state(scene, orc):
    """The reference's steady state with a static camera: CloudVolume::sortBoards has left the
    billboard arrays far->near, and that order is the instance order of both draws."""
    scene.board_pos, scene.board_scale = orc.sort_boards(scene.board_pos, scene.board_scale, scene.vol.position, scene.cam.position)
    return scene
