"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs — never by anything under cloud-renderer_b200/.  The oracle takes the same POD structs
as the C-ABI (include/cloud_renderer_b200.h), so a `Scene` is handed over byte for byte.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None
_scene_cls = None


def build(force=False):
    """(re)builds liboracle.so through the Makefile, which knows the header dependency"""
    subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []) + ["-s", "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_board_distance.restype = C.c_float
    return _lib


class TraceStats(C.Structure):
    _fields_ = [("fragments", C.c_uint64), ("coneSamples", C.c_uint64), ("noiseSamples", C.c_uint64), ("rectPixels", C.c_uint64)]


def _scene_struct(scene):
    global _scene_cls
    if _scene_cls is None:
        class OrcScene(C.Structure):
            _fields_ = [("vol", type(scene.vol)), ("sun", type(scene.sun)), ("cam", type(scene.cam)), ("tp", type(scene.tp)),
                        ("width", C.c_int32), ("height", C.c_int32), ("n_boards", C.c_int32),
                        ("board_pos", C.c_void_p), ("board_scale", C.c_void_p), ("noise", C.c_void_p), ("noise_dim", C.c_int32)]
        _scene_cls = OrcScene
    s = _scene_cls()
    s.vol, s.sun, s.cam, s.tp = scene.vol, scene.sun, scene.cam, scene.tp
    s.width, s.height, s.n_boards = scene.width, scene.height, scene.n_boards
    # keep the arrays alive on the struct
    s._pos = np.ascontiguousarray(scene.board_pos, dtype=np.float32)
    s._scale = np.ascontiguousarray(scene.board_scale, dtype=np.float32)
    s._noise = np.ascontiguousarray(scene.noise, dtype=np.int8)
    s.board_pos, s.board_scale, s.noise = s._pos.ctypes.data, s._scale.ctypes.data, s._noise.ctypes.data
    s.noise_dim = round((s._noise.size // 4) ** (1.0 / 3.0))
    return s


def num_threads():
    return lib().orc_num_threads()


def sun_update(vol, sun, derived_cls):
    out = derived_cls()
    lib().orc_sun_update(C.byref(vol), C.byref(sun), C.byref(out))
    return out


def camera_update(width, height, eye, look_at, camera_cls):
    out = camera_cls()
    e, l = (C.c_float * 3)(*eye), (C.c_float * 3)(*look_at)
    lib().orc_camera_update(width, height, e, l, C.byref(out))
    return out


def sort_boards(pos, scale, volpos, point):
    """CloudVolume::sortBoards: returns sorted copies (far -> near)."""
    p = np.array(pos, dtype=np.float32, order="C", copy=True)
    s = np.array(scale, dtype=np.float32, order="C", copy=True)
    vp, pt = (C.c_float * 3)(*volpos), (C.c_float * 3)(*point)
    lib().orc_sort_boards(C.c_void_p(p.ctypes.data), C.c_void_p(s.ctypes.data), len(s), vp, pt)
    return p, s


def board_distances(pos, volpos, point):
    vp, pt = (C.c_float * 3)(*volpos), (C.c_float * 3)(*point)
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    return np.array([lib().orc_board_distance((C.c_float * 3)(*pos[i]), vp, pt) for i in range(len(pos))], dtype=np.float32)


def build_noise(alpha):
    alpha = np.ascontiguousarray(alpha, dtype=np.int8)
    dim = round(alpha.size ** (1.0 / 3.0))
    out = np.empty((alpha.size, 4), dtype=np.int8)
    lib().orc_build_noise(C.c_void_p(alpha.ctypes.data), dim, C.c_void_p(out.ctypes.data))
    return out


def voxelize(scene, want_posmap=True):
    """-> (posmap (H,W,4) f32, depth (H,W) f32, level0 (D,D,D) u8)"""
    s = _scene_struct(scene)
    D = scene.vol.dimension
    posmap = np.zeros((scene.height, scene.width, 4), dtype=np.float32) if want_posmap else None
    depth = np.zeros((scene.height, scene.width), dtype=np.float32) if want_posmap else None
    level0 = np.zeros((D, D, D), dtype=np.uint8)
    lib().orc_voxelize(C.byref(s), C.c_void_p(posmap.ctypes.data if want_posmap else None),
                       C.c_void_p(depth.ctypes.data if want_posmap else None), C.c_void_p(level0.ctypes.data))
    return posmap, depth, level0


def chain_size(D, levels):
    return sum(max(1, D >> l) ** 3 for l in range(levels))


def mips(level0, levels):
    D = level0.shape[0]
    chain = np.zeros(chain_size(D, levels), dtype=np.uint8)
    l0 = np.ascontiguousarray(level0, dtype=np.uint8)
    lib().orc_mips(C.c_void_p(l0.ctypes.data), D, levels, C.c_void_p(chain.ctypes.data))
    return chain


def mips_f32(level0, levels):
    D = level0.shape[0]
    chain = np.zeros(chain_size(D, levels), dtype=np.float32)
    l0 = np.ascontiguousarray(level0, dtype=np.float32)
    lib().orc_mips_f32(C.c_void_p(l0.ctypes.data), D, levels, C.c_void_p(chain.ctypes.data))
    return chain


def chain_level(chain, D, level):
    off = sum(max(1, D >> l) ** 3 for l in range(level))
    s = max(1, D >> level)
    return chain[off:off + s ** 3].reshape(s, s, s)


def cone_trace(scene, chain, quantize_fb8=False, rows=None, want_u8=True):
    """Draws scene.board_* IN ARRAY ORDER (sort first, as coneTrace does).
    -> (image f32 (H,W,4), image u8 (H,W,4) or None, TraceStats)"""
    s = _scene_struct(scene)
    H, W = scene.height, scene.width
    img = np.zeros((H, W, 4), dtype=np.float32)
    u8 = np.zeros((H, W, 4), dtype=np.uint8) if want_u8 else None
    st = TraceStats()
    r0, r1 = rows if rows else (0, H)
    ch = np.ascontiguousarray(chain, dtype=np.uint8)
    lib().orc_cone_trace(C.byref(s), C.c_void_p(ch.ctypes.data), C.c_void_p(img.ctypes.data),
                         C.c_void_p(u8.ctypes.data if want_u8 else None), 1 if quantize_fb8 else 0, r0, r1, C.byref(st))
    return img, u8, st


def board_rects(scene, which):
    s = _scene_struct(scene)
    out = np.zeros((scene.n_boards, 4), dtype=np.int32)
    r = (C.c_int32 * 4)()
    for b in range(scene.n_boards):
        lib().orc_board_rect(C.byref(s), which, b, r)
        out[b] = r[:]
    return out


def conetrace_fragment(scene, chain, frag_pos, frag_tex, center, radius):
    s = _scene_struct(scene)
    col = (C.c_float * 4)()
    ch = np.ascontiguousarray(chain, dtype=np.uint8)
    ok = lib().orc_conetrace_fragment(C.byref(s), C.c_void_p(ch.ctypes.data), (C.c_float * 3)(*frag_pos), (C.c_float * 2)(*frag_tex),
                                      (C.c_float * 3)(*center), C.c_float(radius), col)
    return bool(ok), np.array(col[:], dtype=np.float32)


def first_voxelize_fragment(scene, frag_pos, center, radius):
    s = _scene_struct(scene)
    wp, d = (C.c_float * 3)(), C.c_float()
    ok = lib().orc_first_voxelize_fragment(C.byref(s), (C.c_float * 3)(*frag_pos), (C.c_float * 3)(*center), C.c_float(radius), wp, C.byref(d))
    return bool(ok), np.array(wp[:], dtype=np.float32), d.value


def second_voxelize_indices(vol, world_pos):
    out = (C.c_int32 * 27)()
    lib().orc_second_voxelize_indices(C.byref(vol), (C.c_float * 3)(*world_pos), out)
    return np.array(out[:], dtype=np.int32).reshape(9, 3)


def cone_lods(tp):
    n = tp.vctSteps
    lods, hs = (C.c_float * n)(), (C.c_float * n)()
    lib().orc_cone_lods(C.byref(tp), lods, hs)
    return np.array(lods[:], dtype=np.float32), np.array(hs[:], dtype=np.float32)
