"""cloud-renderer_b200 — ctypes view of the C-ABI in include/cloud_renderer_b200.h.

The product is the shared library built from csrc/ (hand-written sm_100a CUDA behind a C
boundary).  This module is the thin Python host used by tests/ and bench.py: struct mirrors,
a `Renderer` wrapper with the reference's call shape (set parameters, `voxelize()`,
`cone_trace()`), and nothing else.  There is no CPU path here: if the library is missing or
no GPU is present the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CRN_LIB") or os.path.join(_HERE, "libcloud_renderer_b200.so")      # CRN_LIB: experiment builds (build.py build_variant)

CRN_OK, CRN_ERR_INVALID_ARG, CRN_ERR_CUDA, CRN_ERR_STATE, CRN_ERR_UNSUPPORTED, CRN_ERR_NO_DEVICE = range(6)
MEM_HOST, MEM_DEVICE = 0, 1
IMAGE_RGBA8, IMAGE_RGBA32F = 0, 1
VOLUME_R8, VOLUME_R32F, VOLUME_RG8 = 0, 1, 2
SAMPLER_EXPLICIT, SAMPLER_TEXTURE = 0, 1

f32, i32, u64 = C.c_float, C.c_int32, C.c_uint64


class VolumeDesc(C.Structure):
    _fields_ = [("dimension", i32), ("levels", i32), ("position", f32 * 3), ("xBounds", f32 * 2), ("yBounds", f32 * 2),
                ("zBounds", f32 * 2), ("fluffiness", f32), ("format", i32)]


class Sun(C.Structure):
    _fields_ = [("position", f32 * 3), ("innerColor", f32 * 3), ("outerColor", f32 * 3), ("innerRadius", f32), ("outerRadius", f32)]


class SunDerived(C.Structure):
    _fields_ = [("P", f32 * 16), ("V", f32 * 16), ("nearPlane", f32 * 3), ("farPlane", f32 * 3), ("clipDistance", f32)]


class Camera(C.Structure):
    _fields_ = [("P", f32 * 16), ("V", f32 * 16), ("position", f32 * 3)]


class TraceParams(C.Structure):
    _fields_ = [("stepSize", f32), ("noiseOpacity", f32), ("numOctaves", i32), ("freqStep", f32), ("persStep", f32),
                ("adjustSize", f32), ("minNoiseSteps", i32), ("maxNoiseSteps", i32), ("minNoiseColor", f32),
                ("noiseColorScale", f32), ("windVel", f32 * 3),
                ("vctSteps", i32), ("vctConeAngle", f32), ("vctConeInitialHeight", f32), ("vctLodOffset", f32),
                ("vctDownScaling", f32),
                ("showQuad", i32), ("doConeTrace", i32), ("doNoiseSample", i32),
                ("runTime", f32),
                ("clearColor", f32 * 4), ("drawSun", i32), ("transmittanceCutoff", f32), ("sampler", i32), ("skipEmptySpace", i32),
                ("quantizeFramebuffer", i32)]


class TraceStats(C.Structure):
    _fields_ = [("fragments", u64), ("coneSamples", u64), ("noiseSamples", u64), ("binEntries", u64), ("coneSamplesSkipped", u64), ("filteredFetches", u64),
                ("bakedFetches", u64), ("noiseLatticeSteps", u64), ("codeLookups", u64)]


class Timings(C.Structure):
    _fields_ = [("prepSortMs", f32), ("lightBinMs", f32), ("voxelizeMs", f32), ("mipMs", f32), ("camBinMs", f32), ("traceMs", f32),
                ("coneAccelMs", f32)]


# every symbol include/cloud_renderer_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "crn_create", "crn_destroy", "crn_last_error", "crn_sync", "crn_set_volume", "crn_set_billboards", "crn_set_sun",
    "crn_regenerate_billboards", "crn_animate_billboards", "crn_read_billboards", "crn_read_volume_alpha", "crn_export_voxels",
    "crn_sun_update", "crn_set_camera", "crn_camera_update", "crn_set_window", "crn_set_trace_params",
    "crn_default_trace_params", "crn_set_noise", "crn_build_noise", "crn_voxelize", "crn_cone_trace", "crn_cone_trace_async",
    "crn_wait_images", "crn_cone_trace_enqueue", "crn_image_ptr", "crn_set_row_range",
    "crn_set_tile_row_interleave",
    "crn_set_z_slab", "crn_volume_level_ptr", "crn_volume_bits_ptr", "crn_finish_mips", "crn_read_volume",
    "crn_count_active_voxels", "crn_keep_position_map", "crn_read_position_map", "crn_read_sorted_order", "crn_read_bins",
    "crn_get_trace_stats", "crn_set_stats", "crn_set_timing", "crn_get_timings", "crn_get_launch_count", "crn_version",
    "crn_microbench",
]

_lib = None


def load_library():
    """dlopen the C-ABI library. Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). "
                           "cloud-renderer_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.crn_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    lib.crn_destroy.argtypes = [vp]
    lib.crn_destroy.restype = None
    lib.crn_last_error.argtypes = [vp]
    lib.crn_last_error.restype = C.c_char_p
    lib.crn_version.restype = C.c_char_p
    lib.crn_sync.argtypes = [vp]
    lib.crn_set_volume.argtypes = [vp, C.POINTER(VolumeDesc)]
    lib.crn_set_billboards.argtypes = [vp, vp, vp, i32, i32]
    lib.crn_regenerate_billboards.argtypes = [vp, i32, C.POINTER(f32 * 3), C.POINTER(f32 * 3), f32, f32, C.c_double, C.c_uint64]
    lib.crn_animate_billboards.argtypes = [vp, C.c_double]
    lib.crn_read_billboards.argtypes = [vp, vp, vp]
    lib.crn_set_sun.argtypes = [vp, C.POINTER(Sun)]
    lib.crn_sun_update.argtypes = [C.POINTER(VolumeDesc), C.POINTER(Sun), C.POINTER(SunDerived)]
    lib.crn_set_camera.argtypes = [vp, C.POINTER(Camera)]
    lib.crn_camera_update.argtypes = [i32, i32, C.POINTER(f32 * 3), C.POINTER(f32 * 3), C.POINTER(Camera)]
    lib.crn_set_window.argtypes = [vp, i32, i32]
    lib.crn_set_trace_params.argtypes = [vp, C.POINTER(TraceParams)]
    lib.crn_default_trace_params.argtypes = [C.POINTER(TraceParams)]
    lib.crn_default_trace_params.restype = None
    lib.crn_set_noise.argtypes = [vp, vp, i32]
    lib.crn_build_noise.argtypes = [vp, i32, vp]
    lib.crn_voxelize.argtypes = [vp]
    lib.crn_cone_trace.argtypes = [vp, vp, i32, i32]
    lib.crn_cone_trace_async.argtypes = [vp, vp, i32]
    lib.crn_wait_images.argtypes = [vp]
    lib.crn_cone_trace_enqueue.argtypes = [vp, i32]
    lib.crn_image_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    lib.crn_set_row_range.argtypes = [vp, i32, i32]
    lib.crn_set_z_slab.argtypes = [vp, i32, i32]
    lib.crn_set_tile_row_interleave.argtypes = [vp, i32, i32]
    lib.crn_volume_level_ptr.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(C.c_size_t)]
    lib.crn_volume_bits_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    lib.crn_finish_mips.argtypes = [vp, i32]
    lib.crn_read_volume.argtypes = [vp, i32, vp]
    lib.crn_read_volume_alpha.argtypes = [vp, i32, vp]
    lib.crn_export_voxels.argtypes = [vp, i32, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.crn_count_active_voxels.argtypes = [vp, C.POINTER(u64)]
    lib.crn_keep_position_map.argtypes = [vp, i32]
    lib.crn_read_position_map.argtypes = [vp, vp]
    lib.crn_read_sorted_order.argtypes = [vp, vp]
    lib.crn_read_bins.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), vp, vp, C.POINTER(u64)]
    lib.crn_get_trace_stats.argtypes = [vp, C.POINTER(TraceStats)]
    lib.crn_set_stats.argtypes = [vp, i32]
    lib.crn_set_timing.argtypes = [vp, i32]
    lib.crn_get_timings.argtypes = [vp, C.POINTER(Timings)]
    lib.crn_get_launch_count.argtypes = [vp, C.POINTER(u64)]
    lib.crn_microbench.argtypes = [C.c_int, i32, C.POINTER(C.c_double)]
    _lib = lib
    return lib


class CrnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"crn status {code}: {msg}")
        self.code = code


def default_trace_params():
    p = TraceParams()
    load_library().crn_default_trace_params(C.byref(p))
    return p


def sun_update(vol, sun):
    out = SunDerived()
    load_library().crn_sun_update(C.byref(vol), C.byref(sun), C.byref(out))
    return out


def camera_update(width, height, eye, look_at):
    out = Camera()
    e, l = (f32 * 3)(*eye), (f32 * 3)(*look_at)
    load_library().crn_camera_update(width, height, C.byref(e), C.byref(l), C.byref(out))
    return out


def build_noise(alpha):
    """The deterministic half of ConeTraceShader::initNoiseMap: alpha[dim^3] int8 -> rgba[dim^3,4] int8."""
    alpha = np.ascontiguousarray(alpha, dtype=np.int8)
    dim = round(alpha.size ** (1.0 / 3.0))
    assert dim ** 3 == alpha.size
    out = np.empty((alpha.size, 4), dtype=np.int8)
    rc = load_library().crn_build_noise(alpha.ctypes.data, dim, out.ctypes.data)
    if rc:
        raise CrnError(rc, "crn_build_noise")
    return out


def _ptr(x):
    """host numpy array, or anything with data_ptr() (a torch tensor) -> (address, mem kind)."""
    if isinstance(x, np.ndarray):
        return x.ctypes.data, MEM_HOST
    if hasattr(x, "data_ptr"):
        return x.data_ptr(), (MEM_DEVICE if x.is_cuda else MEM_HOST)
    raise TypeError(type(x))


class Renderer:
    """One context = one (device, stream).  Mirrors the reference's frame:
        Sun::update / volume->update  -> set_*()
        voxelizeShader->voxelize      -> voxelize()
        coneShader->coneTrace         -> cone_trace()
    """

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.crn_create(device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc:
            raise CrnError(rc, self.lib.crn_last_error(None).decode())
        self.h = h
        self.width = self.height = 0
        self.vol = None
        self.n_boards = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.crn_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc:
            raise CrnError(rc, self.lib.crn_last_error(self.h).decode())

    # ---- parameter surface
    def set_volume(self, vol):
        self.vol = vol
        self._ck(self.lib.crn_set_volume(self.h, C.byref(vol)))

    def set_billboards(self, positions, scales):
        if isinstance(positions, np.ndarray):           # host arrays: the C side reads raw float32 memory
            positions = np.ascontiguousarray(positions, dtype=np.float32)
            scales = np.ascontiguousarray(scales, dtype=np.float32)
        else:
            assert positions.is_contiguous() and scales.is_contiguous() and str(positions.dtype).endswith("float32") and str(scales.dtype).endswith("float32")
        n = int(scales.shape[0])
        assert tuple(positions.shape) == (n, 3) and tuple(scales.shape) == (n,), "positions must be (n,3), scales (n,)"
        p, mp = _ptr(positions)
        s, ms = _ptr(scales)
        assert mp == ms
        self._keep = (positions, scales)                # the upload is asynchronous: keep the source alive
        self._ck(self.lib.crn_set_billboards(self.h, p, s, n, mp))
        self.n_boards = n

    def regenerate_billboards(self, count, min_offset, max_offset, min_scale, max_scale, radius_factor=1.0, seed=0):
        """device-side CloudVolume::regenerateBillboards (src/CloudVolume.cpp:120-137); no host->device copy"""
        lo, hi = (f32 * 3)(*min_offset), (f32 * 3)(*max_offset)
        self._ck(self.lib.crn_regenerate_billboards(self.h, count, C.byref(lo), C.byref(hi), min_scale, max_scale,
                                                    radius_factor, seed & ((1 << 64) - 1)))
        self.n_boards = count

    def animate_billboards(self, angle):
        """offsets = R_y(angle) * base offsets, on the device"""
        self._ck(self.lib.crn_animate_billboards(self.h, float(angle)))

    def read_billboards(self, count=None):
        assert count is None or count == self.n_boards, "the C side copies every billboard of the context"
        count = self.n_boards
        pos = np.empty((count, 3), np.float32)
        scale = np.empty(count, np.float32)
        self._ck(self.lib.crn_read_billboards(self.h, pos.ctypes.data, scale.ctypes.data))
        return pos, scale

    def set_sun(self, sun):
        self._ck(self.lib.crn_set_sun(self.h, C.byref(sun)))

    def set_camera(self, cam):
        self._ck(self.lib.crn_set_camera(self.h, C.byref(cam)))

    def set_window(self, width, height):
        self.width, self.height = width, height
        self._ck(self.lib.crn_set_window(self.h, width, height))

    def set_trace_params(self, tp):
        self._ck(self.lib.crn_set_trace_params(self.h, C.byref(tp)))

    def set_noise(self, rgba):
        rgba = np.ascontiguousarray(rgba, dtype=np.int8)
        dim = round((rgba.size // 4) ** (1.0 / 3.0))
        assert dim ** 3 * 4 == rgba.size
        self._ck(self.lib.crn_set_noise(self.h, rgba.ctypes.data, dim))

    def set_scene(self, scene):
        """everything a tests/bench `Scene` carries"""
        self.set_volume(scene.vol)
        self.set_sun(scene.sun)
        self.set_camera(scene.cam)
        self.set_window(scene.width, scene.height)
        self.set_trace_params(scene.tp)
        self.set_noise(scene.noise)
        self.set_billboards(scene.board_pos, scene.board_scale)

    # ---- passes
    def voxelize(self):
        self._ck(self.lib.crn_voxelize(self.h))

    def cone_trace(self, out=None, fmt=IMAGE_RGBA8):
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.uint8 if fmt == IMAGE_RGBA8 else np.float32)
        p, m = _ptr(out)
        self._ck(self.lib.crn_cone_trace(self.h, p, m, fmt))
        return out

    def cone_trace_async(self, out, fmt=IMAGE_RGBA8):
        p, m = _ptr(out)
        assert m == MEM_HOST
        self._ck(self.lib.crn_cone_trace_async(self.h, p, fmt))

    def cone_trace_enqueue(self, fmt=IMAGE_RGBA8):
        """enqueue only; the image stays in the context's device buffer (image_ptr)"""
        self._ck(self.lib.crn_cone_trace_enqueue(self.h, fmt))

    def image_ptr(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.lib.crn_image_ptr(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def wait_images(self):
        self._ck(self.lib.crn_wait_images(self.h))

    def sync(self):
        self._ck(self.lib.crn_sync(self.h))

    # ---- sharding hooks
    def set_row_range(self, r0, r1):
        self._ck(self.lib.crn_set_row_range(self.h, r0, r1))

    def set_tile_row_interleave(self, index, count):
        self._ck(self.lib.crn_set_tile_row_interleave(self.h, index, count))

    def set_z_slab(self, z0, z1):
        self._ck(self.lib.crn_set_z_slab(self.h, z0, z1))

    def volume_level_ptr(self, level):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.lib.crn_volume_level_ptr(self.h, level, C.byref(p), C.byref(n)))
        return p.value, n.value

    def volume_bits_ptr(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.lib.crn_volume_bits_ptr(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def finish_mips(self, first_level):
        self._ck(self.lib.crn_finish_mips(self.h, first_level))

    # ---- inspection
    def read_volume(self, level=0):
        s = max(1, self.vol.dimension >> level)
        out = np.empty((s, s, s), dtype=np.float32 if self.vol.format == VOLUME_R32F else np.uint8)
        self._ck(self.lib.crn_read_volume(self.h, level, out.ctypes.data))
        return out

    def read_chain(self):
        return np.concatenate([self.read_volume(l).ravel() for l in range(self.vol.levels)])

    def read_volume_alpha(self, level=0):
        """CRN_VOLUME_RG8: one level of the occupancy (alpha) channel"""
        s = max(1, self.vol.dimension >> level)
        out = np.empty((s, s, s), dtype=np.uint8)
        self._ck(self.lib.crn_read_volume_alpha(self.h, level, out.ctypes.data))
        return out

    def read_chain_alpha(self):
        return np.concatenate([self.read_volume_alpha(l).ravel() for l in range(self.vol.levels)])

    def export_voxels(self, channel=0):
        """VoxelShader::updateVoxelData: (n,4) float32 rows = world position of each non-empty voxel + its value"""
        n = u64()
        self._ck(self.lib.crn_export_voxels(self.h, channel, None, 0, C.byref(n)))
        out = np.empty((n.value, 4), dtype=np.float32)
        if n.value:
            self._ck(self.lib.crn_export_voxels(self.h, channel, out.ctypes.data, n.value, C.byref(n)))
        return out

    def count_active_voxels(self):
        n = u64()
        self._ck(self.lib.crn_count_active_voxels(self.h, C.byref(n)))
        return n.value

    def keep_position_map(self, on=True):
        self._ck(self.lib.crn_keep_position_map(self.h, 1 if on else 0))

    def read_position_map(self):
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._ck(self.lib.crn_read_position_map(self.h, out.ctypes.data))
        return out

    def read_sorted_order(self, n=None):
        assert n is None or n == self.n_boards, "the C side copies one entry per billboard of the context"
        out = np.empty(self.n_boards, dtype=np.int32)
        self._ck(self.lib.crn_read_sorted_order(self.h, out.ctypes.data))
        return out

    def read_bins(self, which):
        tx, ty, tw, th, tot = i32(), i32(), i32(), i32(), u64()
        self._ck(self.lib.crn_read_bins(self.h, which, C.byref(tx), C.byref(ty), C.byref(tw), C.byref(th), None, None, C.byref(tot)))
        counts = np.empty(tx.value * ty.value, dtype=np.int32)
        entries = np.empty(max(tot.value, 1), dtype=np.int32)
        self._ck(self.lib.crn_read_bins(self.h, which, C.byref(tx), C.byref(ty), C.byref(tw), C.byref(th), counts.ctypes.data,
                                        entries.ctypes.data, C.byref(tot)))
        return dict(tiles_x=tx.value, tiles_y=ty.value, tile_w=tw.value, tile_h=th.value, counts=counts, entries=entries[:tot.value])

    def set_stats(self, on=True):
        self._ck(self.lib.crn_set_stats(self.h, 1 if on else 0))

    def trace_stats(self):
        s = TraceStats()
        self._ck(self.lib.crn_get_trace_stats(self.h, C.byref(s)))
        return s

    def set_timing(self, on=True):
        self._ck(self.lib.crn_set_timing(self.h, 1 if on else 0))

    def timings(self):
        t = Timings()
        self._ck(self.lib.crn_get_timings(self.h, C.byref(t)))
        return t

    def launch_count(self):
        n = u64()
        self._ck(self.lib.crn_get_launch_count(self.h, C.byref(n)))
        return n.value


MICROBENCH = {0: "tex3D trilinear RGBA8_SNORM 32^3 (L1)", 1: "tex3D trilinear R8 256^3 (L2)", 2: "LDG.32 L1-hit",
              3: "global atomicOr (RED), 2 MB set", 4: "shared atomicOr", 5: "FFMA",
              6: "tex2DLayered bilinear RGBA8_SNORM 32x32x32 (L1)",
              7: "tex2DLayered bilinear RG16 UNORM 129x129x128", 8: "tex2DLayered bilinear RG8 UNORM 129x129x128",
              9: "tex2DLayered bilinear RGBA8_SNORM f16x2 return", 10: "tex3D trilinear R16 UNORM 129^3",
              11: "tex3DLod R8 mipmapped 256^3 LOD 4.5", 12: "tex3DLod R8 mipmapped 256^3 LOD 2.5",
              13: "tex2DLayered bilinear RG16F 129x129x128", 14: "tex3D RGBA8_SNORM 32^3, z on slice centres",
              15: "tex2DLayered bilinear RGBA16_SNORM 161x161x160, strided (L1 misses)",
              16: "tex2DLayered bilinear RGBA8_SNORM 161x161x160, strided (L1 misses)",
              17: "tex3D trilinear RG16_SNORM 161^3, strided (L1 misses)",
              18: "tex2DLayered bilinear RGBA16_SNORM 161x161, one layer (L1 hits)"}


def microbench(which, device=0):
    """giga lane-operations per second of one hardware ceiling (see crn_microbench)"""
    g = C.c_double()
    rc = load_library().crn_microbench(device, which, C.byref(g))
    if rc:
        raise CrnError(rc, "crn_microbench")
    return g.value
