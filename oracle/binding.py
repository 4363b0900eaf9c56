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


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


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


def voxelize_paper(scene):
    """paper variant (CRN_VOLUME_RG8): -> (level0 rgb (D,D,D) u8, level0 alpha (D,D,D) u8)"""
    s = _scene_struct(scene)
    D = scene.vol.dimension
    level0 = np.zeros((D, D, D), dtype=np.uint8)
    alpha0 = np.zeros((D, D, D), dtype=np.uint8)
    lib().orc_voxelize_paper(C.byref(s), C.c_void_p(level0.ctypes.data), C.c_void_p(alpha0.ctypes.data))
    return level0, alpha0


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
    ch = np.ascontiguousarray(chain, dtype=np.float32 if scene.vol.format == 1 else np.uint8)     # 1 = CRN_VOLUME_R32F
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
    ch = np.ascontiguousarray(chain, dtype=np.float32 if scene.vol.format == 1 else np.uint8)
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


def list_fragments(scene, which):
    """(n,8) float32: i, j, board, fragPos xyz, fragTex xy — the fragments of the billboard draw, in draw order"""
    s = _scene_struct(scene)
    lib().orc_list_fragments.restype = C.c_int64
    n = lib().orc_list_fragments(C.byref(s), which, C.c_int64(0), None)
    out = np.zeros((n, 8), dtype=np.float32)
    lib().orc_list_fragments(C.byref(s), which, C.c_int64(n), C.c_void_p(out.ctypes.data))
    return out


# ------------------------------------------------------------------------------------------------
# oracle/_ref/libref_glsl.so: the reference's own shaders compiled as C++ (oracle/ref_glsl/build_ref.sh)
# ------------------------------------------------------------------------------------------------
REF_LIB_PATH = os.path.join(_HERE, "_ref", "libref_glsl.so")
_ref = None


class RefConetraceUniforms(C.Structure):
    _fields_ = [("V", C.c_float * 16), ("lightPos", C.c_float * 3), ("showQuad", C.c_int), ("voxelDim", C.c_int),
                ("xBounds", C.c_float * 2), ("yBounds", C.c_float * 2), ("zBounds", C.c_float * 2),
                ("doConeTrace", C.c_int), ("vctSteps", C.c_int), ("vctConeAngle", C.c_float), ("vctConeInitialHeight", C.c_float),
                ("vctLodOffset", C.c_float), ("vctDownScaling", C.c_float),
                ("doNoise", C.c_int), ("octaveOffsets", C.c_float * 3), ("stepSize", C.c_float), ("noiseOpacity", C.c_float),
                ("numOctaves", C.c_int), ("freqStep", C.c_float), ("persStep", C.c_float), ("adjustSize", C.c_float),
                ("minNoiseSteps", C.c_int), ("maxNoiseSteps", C.c_int), ("minNoiseColor", C.c_float), ("noiseColorScale", C.c_float)]


def ref_available():
    return os.path.exists(REF_LIB_PATH)


def ref_lib():
    global _ref
    if _ref is None:
        lib()                                       # liboracle.so first: libref_glsl.so links against it
        _ref = C.CDLL(REF_LIB_PATH)
    return _ref


def _f(*v):
    return (C.c_float * len(v))(*[np.float32(x) for x in v])


def ref_conetrace_uniforms(scene):
    """the uniform uploads of ConeTraceShader::coneTrace / bindVolume (src/Shaders/ConeTraceShader.cpp:26-69,84-93)"""
    u = RefConetraceUniforms()
    tp, vol = scene.tp, scene.vol
    u.V[:] = scene.cam.V[:]
    u.lightPos[:] = scene.sun.position[:]
    u.showQuad, u.voxelDim = tp.showQuad, vol.dimension
    f = np.float32
    u.xBounds[:] = [f(vol.position[0]) + f(vol.xBounds[k]) for k in range(2)]
    u.yBounds[:] = [f(vol.position[1]) + f(vol.yBounds[k]) for k in range(2)]
    u.zBounds[:] = [f(vol.position[2]) + f(vol.zBounds[k]) for k in range(2)]
    u.doConeTrace, u.vctSteps, u.vctConeAngle, u.vctConeInitialHeight = tp.doConeTrace, tp.vctSteps, tp.vctConeAngle, tp.vctConeInitialHeight
    u.vctLodOffset, u.vctDownScaling = tp.vctLodOffset, tp.vctDownScaling
    u.doNoise = tp.doNoiseSample
    u.octaveOffsets[:] = [f(tp.windVel[k]) * f(tp.runTime) for k in range(3)]
    u.stepSize, u.noiseOpacity, u.numOctaves, u.freqStep, u.persStep, u.adjustSize = tp.stepSize, tp.noiseOpacity, tp.numOctaves, tp.freqStep, tp.persStep, tp.adjustSize
    u.minNoiseSteps, u.maxNoiseSteps, u.minNoiseColor, u.noiseColorScale = tp.minNoiseSteps, tp.maxNoiseSteps, tp.minNoiseColor, tp.noiseColorScale
    return u


def ref_conetrace_fragment(scene, uniforms, chain, frag_pos, frag_nor, frag_tex, center, radius):
    col = (C.c_float * 4)()
    ch = np.ascontiguousarray(chain, dtype=np.uint8)
    nz = np.ascontiguousarray(scene.noise, dtype=np.int8)
    ok = ref_lib().ref_conetrace_fragment(C.byref(uniforms), C.c_void_p(ch.ctypes.data), scene.vol.levels, C.c_void_p(nz.ctypes.data),
                                          round((nz.size // 4) ** (1 / 3)), _f(*frag_pos), _f(*frag_nor), _f(*frag_tex), _f(*center),
                                          C.c_float(radius), col)
    return bool(ok), np.array(col[:], dtype=np.float32)


def ref_first_voxelize_fragment(frag_pos, frag_nor, center, radius, near_plane, clip):
    col, d = (C.c_float * 4)(), C.c_float()
    ok = ref_lib().ref_first_voxelize_fragment(_f(*frag_pos), _f(*frag_nor), _f(*center), C.c_float(radius), _f(*near_plane), C.c_float(clip), col, C.byref(d))
    return bool(ok), np.array(col[:], dtype=np.float32), d.value


def ref_second_voxelize_fragment(vol, texel):
    f = np.float32
    xb = [f(vol.position[0]) + f(vol.xBounds[k]) for k in range(2)]
    yb = [f(vol.position[1]) + f(vol.yBounds[k]) for k in range(2)]
    zb = [f(vol.position[2]) + f(vol.zBounds[k]) for k in range(2)]
    d = f(vol.dimension)
    step = min(f(f(vol.xBounds[1]) - f(vol.xBounds[0])) / d, f(f(vol.yBounds[1]) - f(vol.yBounds[0])) / d, f(f(vol.zBounds[1]) - f(vol.zBounds[0])) / d)
    idx, val = (C.c_int * 27)(), (C.c_float * 9)()
    n = ref_lib().ref_second_voxelize_fragment(_f(*texel), vol.dimension, _f(*xb), _f(*yb), _f(*zb), C.c_float(step), idx, val)
    return n, np.array(idx[:], dtype=np.int32).reshape(9, 3), np.array(val[:], dtype=np.float32)


def ref_sun_fragment(sun, frag_pos):
    col = (C.c_float * 4)()
    ok = ref_lib().ref_sun_fragment(_f(*frag_pos), _f(*sun.position), _f(*sun.innerColor), _f(*sun.outerColor), C.c_float(sun.innerRadius),
                                    C.c_float(sun.outerRadius), col)
    return bool(ok), np.array(col[:], dtype=np.float32)


def ref_billboard_vertex(P, V, volume_position, vert, board_position, board_scale):
    """billboard_vert_instanced.glsl for one vertex; Vi = transpose(V with zeroed translation) as the drivers build it"""
    Vm = np.array(V[:], dtype=np.float32).reshape(4, 4).copy()      # [col][row]
    Vm[3, :3] = 0.0
    Vi = Vm.T.copy()
    glpos, fpos, fnor, ftex, cen, sc = (C.c_float * 4)(), (C.c_float * 3)(), (C.c_float * 3)(), (C.c_float * 2)(), (C.c_float * 3)(), C.c_float()
    ref_lib().ref_billboard_vertex(_f(*P[:]), _f(*V[:]), _f(*Vi.ravel()), _f(*volume_position), _f(vert[0], vert[1], 0.0), _f(0.0, 0.0, 1.0),
                                   _f(*board_position), C.c_float(board_scale), glpos, fpos, fnor, ftex, cen, C.byref(sc))
    return dict(gl_Position=np.array(glpos[:], np.float32), fragPos=np.array(fpos[:], np.float32), fragNor=np.array(fnor[:], np.float32),
                fragTex=np.array(ftex[:], np.float32), center=np.array(cen[:], np.float32), scale=sc.value)


# ---- the paper variant's interior march (oracle) and the reference shader with the march switched back on -------------
def first_voxelize_march(scene, frag_pos, center, radius, cap=1024):
    """-> (n, idx[n,3]) in-range image stores of the march for one fragment, n = -1 on discard"""
    s = _scene_struct(scene)
    idx = (C.c_int32 * (3 * cap))()
    n = lib().orc_first_voxelize_march(C.byref(s), _f(*frag_pos), _f(*center), C.c_float(radius), idx, cap)
    return n, np.array(idx[:3 * max(0, min(n, cap))], dtype=np.int32).reshape(-1, 3)


def ref_first_voxelize_paper_fragment(vol, frag_pos, frag_nor, center, radius, near_plane, clip, cap=1024):
    f = np.float32
    xb = [f(vol.position[0]) + f(vol.xBounds[k]) for k in range(2)]
    yb = [f(vol.position[1]) + f(vol.yBounds[k]) for k in range(2)]
    zb = [f(vol.position[2]) + f(vol.zBounds[k]) for k in range(2)]
    d = f(vol.dimension)
    step = min(f(f(vol.xBounds[1]) - f(vol.xBounds[0])) / d, f(f(vol.yBounds[1]) - f(vol.yBounds[0])) / d, f(f(vol.zBounds[1]) - f(vol.zBounds[0])) / d)
    idx, col, dep = (C.c_int * (3 * cap))(), (C.c_float * 4)(), C.c_float()
    n = ref_lib().ref_first_voxelize_paper_fragment(_f(*frag_pos), _f(*frag_nor), _f(*center), C.c_float(radius), _f(*near_plane), C.c_float(clip),
                                                    vol.dimension, _f(*xb), _f(*yb), _f(*zb), C.c_float(step), idx, cap, col, C.byref(dep))
    return n, np.array(idx[:3 * max(0, min(n, cap))], dtype=np.int32).reshape(-1, 3), np.array(col[:], dtype=np.float32), dep.value


# ---- the reference's HOST code compiled from /root/reference/src (oracle/ref_glsl/ref_host.cpp) -----------------------
def ref_host_sun_update(volpos, xb, yb, zb, sunpos):
    V, P, n, fa, clip = (C.c_float * 16)(), (C.c_float * 16)(), (C.c_float * 3)(), (C.c_float * 3)(), C.c_float()
    ref_lib().ref_host_sun_update(_f(*volpos), _f(*xb), _f(*yb), _f(*zb), _f(*sunpos), V, P, n, fa, C.byref(clip))
    return dict(V=np.array(V[:], np.float32), P=np.array(P[:], np.float32), nearPlane=np.array(n[:], np.float32),
                farPlane=np.array(fa[:], np.float32), clipDistance=np.float32(clip.value))


def ref_host_sun_defaults():
    p, i, o, ir, orr = (C.c_float * 3)(), (C.c_float * 3)(), (C.c_float * 3)(), C.c_float(), C.c_float()
    ref_lib().ref_host_sun_defaults(p, i, o, C.byref(ir), C.byref(orr))
    return dict(position=np.array(p[:], np.float32), innerColor=np.array(i[:], np.float32), outerColor=np.array(o[:], np.float32),
                innerRadius=ir.value, outerRadius=orr.value)


def ref_host_camera_update(width, height, position, phi, theta):
    P, V, l = (C.c_float * 16)(), (C.c_float * 16)(), (C.c_float * 3)()
    ref_lib().ref_host_camera_update(int(width), int(height), _f(*position), C.c_double(phi), C.c_double(theta), P, V, l)
    return dict(P=np.array(P[:], np.float32), V=np.array(V[:], np.float32), lookAt=np.array(l[:], np.float32))


def ref_host_sort_boards(pos, scale, volpos, point):
    p = np.array(pos, dtype=np.float32, order="C", copy=True)
    s = np.array(scale, dtype=np.float32, order="C", copy=True)
    ref_lib().ref_host_sort_boards(C.c_void_p(p.ctypes.data), C.c_void_p(s.ctypes.data), len(s), _f(*volpos), _f(*point))
    return p, s


def ref_host_voxel_index(dim, volpos, xb, yb, zb, index):
    ijk, w, vs = (C.c_int * 3)(), (C.c_float * 3)(), (C.c_float * 3)()
    ref_lib().ref_host_voxel_index(int(dim), _f(*volpos), _f(*xb), _f(*yb), _f(*zb), int(index), ijk, w, vs)
    return np.array(ijk[:], np.int32), np.array(w[:], np.float32), np.array(vs[:], np.float32)


def ref_host_noise_normals(alpha):
    """ConeTraceShader::initNoiseMap's normal loop on a texture with the given alpha channel -> rgba[dim^3,4] int8"""
    alpha = np.ascontiguousarray(alpha, dtype=np.int8)
    dim = round(alpha.size ** (1.0 / 3.0))
    rgba = np.zeros((alpha.size, 4), dtype=np.int8)
    rgba[:, 3] = alpha
    ref_lib().ref_host_noise_normals(C.c_void_p(rgba.ctypes.data), dim)
    return rgba
