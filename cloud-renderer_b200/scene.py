"""Fixed-seed scene fixtures for the BASELINE.json configs.

The reference seeds rand() with time(0) (src/main.cpp:70), so its scenes are not
reproducible; here every input blob (billboard offsets/scales, the noise texture's alpha
channel) comes from a splitmix64 stream, and both the CUDA library and the oracle are fed
the identical bytes.  Distributions follow the reference's generator:
  billboards  CloudVolume::regenerateBillboards(200, vec3(-2.5), vec3(2.5), 1, 2.5)   src/main.cpp:82
  noise alpha (char)Util::genRandom(-128, 128)                                         src/Shaders/ConeTraceShader.cpp:131-133
Scene constants are the reference's defaults (src/main.cpp:25-43, src/Camera.cpp:8-29).
"""
import math
from dataclasses import dataclass, field

import numpy as np

from . import Camera, Sun, TraceParams, VolumeDesc
from . import build_noise as _lib_build_noise, camera_update as _lib_camera_update, default_trace_params as _lib_default_trace_params

SEED = 0xC10D5EED
_M64 = (1 << 64) - 1


def splitmix64(seed, n):
    """n uint64 values of the splitmix64 stream that starts at `seed` (vectorised)."""
    i = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed & _M64) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform01(seed, n):
    """float64 in [0,1) from the top 53 bits"""
    return (splitmix64(seed, n) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


@dataclass
class Scene:
    name: str
    vol: VolumeDesc
    sun: Sun
    cam: Camera
    tp: TraceParams
    width: int
    height: int
    board_pos: np.ndarray          # (N,3) float32 offsets relative to vol.position
    board_scale: np.ndarray        # (N,)  float32
    noise: np.ndarray              # (dim^3, 4) int8, RGBA8_SNORM texels, x fastest
    eye: tuple = (0.0, 0.0, 0.0)
    look_at: tuple = (1.0, 0.0, 0.0)
    meta: dict = field(default_factory=dict)

    @property
    def n_boards(self):
        return int(self.board_scale.shape[0])


# name: (D, L, N, W, H)
CONFIGS = {
    "tiny": (32, 4, 24, 160, 96),            # seconds on the oracle; parity workhorse
    "small": (64, 5, 120, 320, 192),
    "C1": (32, 4, 200, 1280, 720),           # reference default scene (src/main.cpp:25-33,82)
    "C2": (128, 8, 4000, 1920, 1080),
    "C3": (256, 9, 20000, 3840, 2160),
    "C4": (512, 10, 20000, 7680, 4320),
    "C5": (256, 9, 20000, 3840, 2160),       # one view of the 64-view orbit
}


class LibraryHost:
    """host math of the fixtures from the product library (crn_camera_update, crn_build_noise, crn_default_trace_params)"""
    camera_update = staticmethod(_lib_camera_update)
    build_noise = staticmethod(_lib_build_noise)
    default_trace_params = staticmethod(_lib_default_trace_params)


def reference_default_trace_params():
    """src/Shaders/ConeTraceShader.hpp:15-36 + src/main.cpp:112, without touching the library (bench.py --impl reference)"""
    p = TraceParams()
    p.stepSize, p.noiseOpacity, p.numOctaves, p.freqStep, p.persStep = 0.01, 4.0, 4, 3.0, 0.5
    p.adjustSize, p.minNoiseSteps, p.maxNoiseSteps, p.minNoiseColor, p.noiseColorScale = 40.0, 2, 8, 0.2, 0.45
    p.windVel[:] = (0.01, 0.0, 0.0)
    p.vctSteps, p.vctConeAngle, p.vctConeInitialHeight, p.vctLodOffset, p.vctDownScaling = 16, 0.9, 0.1, 0.0, 1.0
    p.showQuad, p.doConeTrace, p.doNoiseSample = 0, 1, 1
    p.runTime = 0.0
    p.clearColor[:] = (0.2, 0.3, 0.5, 1.0)
    p.drawSun, p.transmittanceCutoff, p.sampler, p.skipEmptySpace, p.quantizeFramebuffer = 1, 0.0, 0, 1, 0
    return p


class OracleHost:
    """the same fixtures from the oracle's host math: the reference arm of bench.py must not load the product library"""

    def __init__(self, orc):
        self.orc = orc

    def camera_update(self, width, height, eye, look_at):
        return self.orc.camera_update(width, height, eye, look_at, Camera)

    def build_noise(self, alpha):
        return self.orc.build_noise(alpha)

    default_trace_params = staticmethod(reference_default_trace_params)


def make_noise(seed=SEED, dim=32, host=LibraryHost):
    a = np.floor(uniform01(seed ^ 0x5EED0001, dim ** 3) * 256.0 - 128.0)     # trunc of U[-128,128)
    alpha = np.clip(a, -128, 127).astype(np.int8)
    return host.build_noise(alpha)


def make_boards(n, seed=SEED, radius_mode="auto"):
    """offsets U[-2.5,2.5]^3, scale U[1,2.5]; for N > 200 the radii shrink by (200/N)^(1/3) so that
    the cloud's fill (sum of sphere volumes) stays what the reference scene has ("fill" mode,
    SURVEY.md §8d); "reference" keeps U[1,2.5] whatever N is."""
    u = uniform01(seed ^ 0xB0A2D5, 4 * n).reshape(n, 4)
    pos = (u[:, :3] * 5.0 - 2.5).astype(np.float32)
    scale = u[:, 3] * 1.5 + 1.0
    if radius_mode == "auto":
        radius_mode = "fill" if n > 200 else "reference"
    if radius_mode == "fill":
        scale = scale * (200.0 / n) ** (1.0 / 3.0)
    return np.ascontiguousarray(pos), np.ascontiguousarray(scale.astype(np.float32)), radius_mode


def animate(pos0, frame, rate=0.2, fps=60.0):
    """C3's animation: rigid rotation of the billboard offsets about +Y at `rate` rad/s."""
    a = rate * frame / fps
    c, s = math.cos(a), math.sin(a)
    out = pos0.copy()
    out[:, 0] = (c * pos0[:, 0] + s * pos0[:, 2]).astype(np.float32)
    out[:, 2] = (-s * pos0[:, 0] + c * pos0[:, 2]).astype(np.float32)
    return out


def make_scene(name="C1", seed=SEED, frame=0, view=None, n_views=64, radius_mode="auto", size=None, boards=None,
               cutoff=0.0, host=LibraryHost):
    D, L, N, W, H = CONFIGS[name]
    if size is not None:
        W, H = size
    if boards is not None:
        N = boards
    vol = VolumeDesc()
    vol.dimension, vol.levels = D, L
    vol.position[:] = (25.0, 0.0, 0.0)
    vol.xBounds[:] = vol.yBounds[:] = vol.zBounds[:] = (-5.0, 5.0)
    vol.fluffiness = 1.0
    vol.format = 0
    sun = Sun()
    sun_pos = np.array([5.0, 20.0, -5.0])
    sun.innerColor[:] = (1.0, 1.0, 1.0)
    sun.outerColor[:] = (1.0, 1.0, 0.0)
    sun.innerRadius, sun.outerRadius = 1.0, 2.0
    eye, look = (0.0, 0.0, 0.0), (1.0, 0.0, 0.0)
    if view is not None:                     # C5: camera orbit + sun azimuth stepping around the volume centre
        ang = 2.0 * math.pi * view / n_views
        centre = np.array([25.0, 0.0, 0.0])
        eye = tuple(centre + 25.0 * np.array([-math.cos(ang), 0.0, math.sin(ang)]))
        look = tuple(centre)
        rel = sun_pos - centre
        c, s = math.cos(ang), math.sin(ang)
        sun_pos = centre + np.array([c * rel[0] + s * rel[2], rel[1], -s * rel[0] + c * rel[2]])
    sun.position[:] = tuple(float(np.float32(v)) for v in sun_pos)
    cam = host.camera_update(W, H, eye, look)
    tp = host.default_trace_params()
    tp.runTime = frame / 60.0
    tp.transmittanceCutoff = cutoff
    pos0, scale, mode = make_boards(N, seed, radius_mode)
    pos = animate(pos0, frame) if frame else pos0
    return Scene(name, vol, sun, cam, tp, W, H, pos, scale, make_noise(seed, host=host), eye, look,
                 dict(radius_mode=mode, frame=frame, view=view, seed=seed))
