"""b200rt — Python binding of libb200rt.so (the C ABI in include/b200rt.h).

The directory name contains hyphens, so it is loaded through ``__graft_entry__.load_package()``,
which registers it as the module ``b200rt``.  The binding mirrors the reference's builder surface
(``World.push_object(...).push_triangles(...)``, ``Camera``, ``load_obj``; main.rs:161-178,
705-746, 778-807) on top of the C entry points, and never computes anything itself: every render
call goes to the CUDA library.  If the library is missing or no GPU is present the calls raise —
there is no CPU fallback (the CPU oracle lives under oracle/ and is test infrastructure only).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Optional, Sequence

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200RT_LIB") or os.path.join(_PKG_DIR, "_lib", "libb200rt.so")   # B200RT_LIB: tuning builds

# ---- error codes / enums (include/b200rt.h) ----------------------------------------------------------
OK, ERR_INVALID, ERR_CUDA, ERR_NO_SCENE, ERR_NO_DEVICE, ERR_IO, ERR_UNSUPPORTED, ERR_NCCL = 0, -1, -2, -3, -4, -5, -6, -7
GROUP_ID_BYTES = 128
FACE_FRONT, FACE_BACK, FACE_BOTH = 0, 1, 2
MATERIAL_COLOR, MATERIAL_GENERATIVE = 0, 1
DIFFUSE_CONST, DIFFUSE_STRIPE_V, DIFFUSE_CHECKER_UPV = 0, 1, 2
NORMAL_CONST, NORMAL_SINCOS_U = 0, 1
LIGHT_DIRECTIONAL, LIGHT_SPOT, LIGHT_POINT = 0, 1, 2
CAST_TWO_PHASE, CAST_BRUTE_EXACT, CAST_BVH = 0, 1, 2
TRACER_WAVEFRONT, TRACER_MEGAKERNEL = 0, 1
OBJ_USE_TEXCOORDS, OBJ_USE_NORMALS = 1, 2
MAX_DEPTH = 16


class B200rtError(RuntimeError):
    def __init__(self, code: int, what: str, detail: str = ""):
        self.code = code
        super().__init__(f"{what}: error {code}" + (f" ({detail})" if detail else ""))


# ---- PODs ------------------------------------------------------------------------------------------
class Vertex(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("normal", C.c_float * 3), ("uv", C.c_float * 2)]


class Triangle(C.Structure):
    _fields_ = [("vertices", Vertex * 3), ("object_index", C.c_uint32)]


class Sphere(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("radius", C.c_float), ("object_index", C.c_uint32)]


class Material(C.Structure):
    _fields_ = [
        ("kind", C.c_uint32), ("normal", C.c_float * 3), ("diffuse_color", C.c_float * 3), ("shiness", C.c_float),
        ("specular_color", C.c_float * 3), ("smoothness", C.c_float), ("transparency", C.c_float),
        ("refraction_index", C.c_float), ("opaque_decay", C.c_float), ("diffuse_fn", C.c_uint32),
        ("normal_fn", C.c_uint32), ("fn_params", C.c_float * 8),
    ]


class Light(C.Structure):
    _fields_ = [
        ("kind", C.c_uint32), ("has_origin", C.c_uint32), ("origin", C.c_float * 3), ("direction", C.c_float * 3),
        ("angle", C.c_float), ("softness", C.c_float), ("color", C.c_float * 3),
    ]


class Scene(C.Structure):
    _fields_ = [
        ("triangles", C.POINTER(Triangle)), ("n_triangles", C.c_uint32),
        ("spheres", C.POINTER(Sphere)), ("n_spheres", C.c_uint32),
        ("materials", C.POINTER(Material)), ("n_materials", C.c_uint32),
        ("lights", C.POINTER(Light)), ("n_lights", C.c_uint32),
    ]


class Camera(C.Structure):
    _fields_ = [("fovy", C.c_float), ("center", C.c_float * 3), ("toward", C.c_float * 3), ("up", C.c_float * 3),
                ("near", C.c_float)]


class Ray(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("direction", C.c_float * 3), ("face_direction", C.c_uint32),
                ("exclude_prim", C.c_int32), ("exclude_face", C.c_uint32)]


class Hit(C.Structure):
    _fields_ = [("prim_id", C.c_int32), ("object_index", C.c_uint32), ("face_direction", C.c_uint32),
                ("distance", C.c_float), ("position", C.c_float * 3), ("normal", C.c_float * 3), ("uv", C.c_float * 2)]


class Params(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("row_begin", C.c_uint32), ("row_count", C.c_uint32),
        ("depth", C.c_int32), ("threshold", C.c_float), ("refract_max_distance", C.c_float),
        ("tir_retries", C.c_uint32), ("focus", C.c_float), ("blur", C.c_float), ("seed", C.c_uint64),
        ("cast_mode", C.c_uint32), ("tracer", C.c_uint32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("casts", C.c_uint64), ("tri_pair_tests", C.c_uint64), ("sph_pair_tests", C.c_uint64),
        ("exact_confirms", C.c_uint64), ("samples", C.c_uint64), ("kernel_ms", C.c_float), ("h2d_ms", C.c_float),
        ("d2h_ms", C.c_float), ("wavefront_rounds", C.c_uint32), ("certify_fallbacks", C.c_uint64),
        ("cast_kernel_ms", C.c_float), ("logic_kernel_ms", C.c_float), ("cast_kernel_launches", C.c_uint32),
        ("kernel_launches", C.c_uint32), ("primary_kernel_ms", C.c_float), ("reserved", C.c_uint32),
    ]


RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("direction", "<f4", 3), ("face_direction", "<u4"),
                      ("exclude_prim", "<i4"), ("exclude_face", "<u4")])
HIT_DTYPE = np.dtype([("prim_id", "<i4"), ("object_index", "<u4"), ("face_direction", "<u4"), ("distance", "<f4"),
                      ("position", "<f4", 3), ("normal", "<f4", 3), ("uv", "<f4", 2)])
assert RAY_DTYPE.itemsize == C.sizeof(Ray) and HIT_DTYPE.itemsize == C.sizeof(Hit)

# include/b200rt_dev.h: development micro-benchmarks (not part of the product ABI)
DEV_SYMBOLS = ["b200rt_filter_bench", "b200rt_pipe_bench", "b200rt_dev_build_bvh", "b200rt_dev_color_pow"]

# every symbol include/b200rt.h declares
EXPORTED_SYMBOLS = [
    "b200rt_create", "b200rt_destroy", "b200rt_strerror", "b200rt_last_cuda_error", "b200rt_device_info",
    "b200rt_upload_scene", "b200rt_render_whitted", "b200rt_render_whitted_device", "b200rt_render_distributed",
    "b200rt_render_distributed_device", "b200rt_render_distributed_strips_device", "b200rt_resolve_device", "b200rt_intersect", "b200rt_intersect_device",
    "b200rt_post_process", "b200rt_post_process_device", "b200rt_encode_srgb8", "b200rt_encode_srgb8_device",
    "b200rt_write_png_rgb8",
    "b200rt_get_stats", "b200rt_reset_stats", "b200rt_set_kernel_timing", "b200rt_measure_fp32_peak",
    "b200rt_group_create", "b200rt_group_unique_id", "b200rt_group_create_rank", "b200rt_group_destroy", "b200rt_group_size",
    "b200rt_group_ctx", "b200rt_group_last_error", "b200rt_group_upload_scene", "b200rt_group_render_distributed",
    "b200rt_group_render_distributed_device", "b200rt_group_render_distributed_rows",
    "b200rt_group_render_distributed_rows_device", "b200rt_group_render_whitted", "b200rt_group_render_whitted_device",
    "b200rt_group_last_render_ms",
    "b200rt_world_new", "b200rt_world_free",
    "b200rt_world_push_object", "b200rt_world_push_triangle", "b200rt_world_push_flat_triangle",
    "b200rt_world_push_square", "b200rt_world_push_sphere", "b200rt_world_push_light", "b200rt_world_load_obj",
    "b200rt_world_load_obj_ex", "b200rt_obj_model_count",
    "b200rt_world_scene", "b200rt_world_fixture", "b200rt_fixture_camera", "b200rt_default_params",
]

_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """dlopen libb200rt.so (built in-tree by build.py).  Raises if it is absent: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the render core is CUDA-only; there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32p, f32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float)
    sig = {
        "b200rt_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "b200rt_destroy": (C.c_int, [vp]),
        "b200rt_strerror": (C.c_char_p, [C.c_int]),
        "b200rt_last_cuda_error": (C.c_char_p, [vp]),
        "b200rt_device_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
        "b200rt_upload_scene": (C.c_int, [vp, C.POINTER(Scene)]),
        "b200rt_render_whitted": (C.c_int, [vp, C.POINTER(Camera), C.POINTER(Params), vp, vp]),
        "b200rt_render_whitted_device": (C.c_int, [vp, C.POINTER(Camera), C.POINTER(Params), vp, vp, vp]),
        "b200rt_render_distributed": (C.c_int, [vp, C.POINTER(Camera), C.POINTER(Params), C.c_uint32, C.c_uint32, vp]),
        "b200rt_render_distributed_device": (C.c_int, [vp, C.POINTER(Camera), C.POINTER(Params), C.c_uint32,
                                                       C.c_uint32, vp, vp]),
        "b200rt_render_distributed_strips_device": (C.c_int, [vp, C.POINTER(Camera), C.POINTER(Params), C.c_uint32, C.c_uint32, vp, vp,
                                                              C.c_uint32, C.c_uint32, C.c_uint32]),
        "b200rt_resolve_device": (C.c_int, [vp, vp, vp, C.c_size_t, vp]),
        "b200rt_intersect": (C.c_int, [vp, vp, C.c_size_t, C.c_uint32, vp]),
        "b200rt_intersect_device": (C.c_int, [vp, vp, C.c_size_t, C.c_uint32, vp, vp]),
        "b200rt_post_process": (C.c_int, [vp, vp, C.c_size_t, C.POINTER(C.c_float)]),
        "b200rt_post_process_device": (C.c_int, [vp, vp, C.c_size_t, vp, vp]),
        "b200rt_encode_srgb8": (C.c_int, [vp, vp, C.c_size_t, vp]),
        "b200rt_encode_srgb8_device": (C.c_int, [vp, vp, C.c_size_t, vp, vp]),
        "b200rt_write_png_rgb8": (C.c_int, [C.c_char_p, vp, C.c_uint32, C.c_uint32]),
        "b200rt_get_stats": (C.c_int, [vp, C.POINTER(Stats)]),
        "b200rt_reset_stats": (C.c_int, [vp]),
        "b200rt_set_kernel_timing": (C.c_int, [vp, C.c_int]),
        "b200rt_measure_fp32_peak": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "b200rt_filter_bench": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]),
        "b200rt_pipe_bench": (C.c_int, [vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_double)]),
        "b200rt_dev_color_pow": (C.c_int, [vp, f32p, f32p, f32p, C.c_size_t]),
        "b200rt_dev_build_bvh": (C.c_int, [C.POINTER(Scene), C.c_int, f32p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                           C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
        "b200rt_group_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]),
        "b200rt_group_unique_id": (C.c_int, [vp, C.c_size_t]),
        "b200rt_group_create_rank": (C.c_int, [C.c_int, C.c_int, C.c_int, vp, C.c_size_t, C.POINTER(vp)]),
        "b200rt_group_destroy": (C.c_int, [vp]),
        "b200rt_group_size": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "b200rt_group_ctx": (C.c_int, [vp, C.c_int, C.POINTER(vp)]),
        "b200rt_group_last_error": (C.c_char_p, [vp]),
        "b200rt_group_upload_scene": (C.c_int, [vp, C.POINTER(Scene)]),
        "b200rt_group_render_distributed": (C.c_int, [vp, C.POINTER(Camera), C.POINTER(Params), C.c_uint32, C.c_uint32, vp]),
        "b200rt_group_render_distributed_device": (C.c_int, [vp, C.POINTER(Camera), C.POINTER(Params), C.c_uint32, C.c_uint32, vp]),
        "b200rt_group_render_distributed_rows": (C.c_int, [vp, C.POINTER(Camera), C.POINTER(Params), C.c_uint32, C.c_uint32, vp]),
        "b200rt_group_render_distributed_rows_device": (C.c_int, [vp, C.POINTER(Camera), C.POINTER(Params), C.c_uint32, C.c_uint32, vp]),
        "b200rt_group_render_whitted": (C.c_int, [vp, C.POINTER(Camera), C.POINTER(Params), vp, vp, C.c_int]),
        "b200rt_group_render_whitted_device": (C.c_int, [vp, C.POINTER(Camera), C.POINTER(Params), vp]),
        "b200rt_group_last_render_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
        "b200rt_world_new": (vp, []),
        "b200rt_world_free": (None, [vp]),
        "b200rt_world_push_object": (C.c_int, [vp, C.POINTER(Material)]),
        "b200rt_world_push_triangle": (C.c_int, [vp, C.c_uint32, C.POINTER(Vertex)]),
        "b200rt_world_push_flat_triangle": (C.c_int, [vp, C.c_uint32, f32p, f32p]),
        "b200rt_world_push_square": (C.c_int, [vp, C.c_uint32, f32p, f32p]),
        "b200rt_world_push_sphere": (C.c_int, [vp, C.c_uint32, f32p, C.c_float]),
        "b200rt_world_push_light": (C.c_int, [vp, C.POINTER(Light)]),
        "b200rt_world_load_obj": (C.c_int, [vp, C.c_uint32, C.c_char_p, C.c_float, f32p]),
        "b200rt_world_load_obj_ex": (C.c_int, [vp, C.c_uint32, C.c_char_p, C.c_float, f32p, C.c_int32, C.c_uint32]),
        "b200rt_obj_model_count": (C.c_int, [C.c_char_p]),
        "b200rt_world_scene": (C.c_int, [vp, C.POINTER(Scene)]),
        "b200rt_world_fixture": (C.c_int, [vp, C.c_char_p]),
        "b200rt_fixture_camera": (None, [C.POINTER(Camera)]),
        "b200rt_default_params": (None, [C.POINTER(Params)]),
    }
    for name, (res, args) in sig.items():
        if not hasattr(lib, name) and name.startswith("b200rt_group_") and os.environ.get("B200RT_LIB"):
            continue                                   # a tuning build of an older ABI (B200RT_LIB): no device groups
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def write_png(path: str, rgb8: np.ndarray) -> None:
    """write_to_file, main.rs:764-776: [h][w][3] u8 -> RGB8 PNG via tmp.png + rename (host only)."""
    img = np.ascontiguousarray(rgb8, dtype=np.uint8)
    if img.ndim != 3 or img.shape[2] != 3:
        raise ValueError("write_png expects [height][width][3] u8")
    _check(load_library().b200rt_write_png_rgb8(os.fsencode(path), img.ctypes.data, img.shape[1], img.shape[0]), "write_png")


def strerror(code: int) -> str:
    return load_library().b200rt_strerror(code).decode()


def _check(code: int, what: str, ctx: Optional["Context"] = None) -> int:
    if code < 0:
        detail = strerror(code)
        if ctx is not None and code == ERR_CUDA:
            detail += ": " + load_library().b200rt_last_cuda_error(ctx._h).decode()
        raise B200rtError(code, what, detail)
    return code


def _f32(a, n) -> "C.Array":
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))
    if arr.size != n:
        raise ValueError(f"expected {n} floats, got {arr.size}")
    return (C.c_float * n)(*arr.tolist())


# ---- builders mirroring the reference ----------------------------------------------------------------
def color_material(diffuse_color=(1.0, 1.0, 1.0), shiness=0.0, specular_color=(1.0, 1.0, 1.0), smoothness=1.0,
                   refraction_index=1.0, opaque_decay=0.0, transparency=0.0, normal=(0.0, 0.0, 1.0)) -> Material:
    """ColorMaterial { .. }  (materials.rs:20-31)"""
    m = Material()
    m.kind = MATERIAL_COLOR
    m.normal[:] = normal
    m.diffuse_color[:] = diffuse_color
    m.shiness = shiness
    m.specular_color[:] = specular_color
    m.smoothness = smoothness
    m.transparency = transparency
    m.refraction_index = refraction_index
    m.opaque_decay = opaque_decay
    m.diffuse_fn = DIFFUSE_CONST
    m.normal_fn = NORMAL_CONST
    return m


def generative_material(diffuse_fn=DIFFUSE_CONST, normal_fn=NORMAL_CONST, freq=1.0, c0=(1.0, 1.0, 1.0),
                        c1=(0.0, 0.0, 0.0), nfreq=1.0, **kw) -> Material:
    """GenerativeMaterial { diffuse_fn, normal_fn, .. }  (materials.rs:70-83) with enumerated closures."""
    m = color_material(**kw)
    m.kind = MATERIAL_GENERATIVE
    m.diffuse_fn = diffuse_fn
    m.normal_fn = normal_fn
    m.fn_params[0] = freq
    m.fn_params[1:4] = c0
    m.fn_params[4:7] = c1
    m.fn_params[7] = nfreq
    return m


def directional_light(direction, color, origin=None) -> Light:
    l = Light()
    l.kind = LIGHT_DIRECTIONAL
    l.has_origin = 0 if origin is None else 1
    if origin is not None:
        l.origin[:] = origin
    l.direction[:] = direction
    l.color[:] = color
    return l


def spot_light(origin, direction, angle_rad, softness, color) -> Light:
    l = Light()
    l.kind = LIGHT_SPOT
    l.has_origin = 1
    l.origin[:] = origin
    l.direction[:] = direction
    l.angle = angle_rad
    l.softness = softness
    l.color[:] = color
    return l


def point_light(origin, color) -> Light:
    l = Light()
    l.kind = LIGHT_POINT
    l.has_origin = 1
    l.origin[:] = origin
    l.color[:] = color
    return l


class ObjectProxy:
    """ObjectProxy (main.rs:700-728): pushes primitives tagged with one object index."""

    def __init__(self, world: "World", object_index: int):
        self.world = world
        self.object_index = object_index

    def push_triangle(self, vertices: Sequence[Vertex]) -> "ObjectProxy":
        arr = (Vertex * 3)(*vertices)
        _check(load_library().b200rt_world_push_triangle(self.world._h, self.object_index, arr), "push_triangle")
        return self

    def push_flat_triangle(self, positions, uvs=((0, 0), (0, 0), (0, 0))) -> "ObjectProxy":
        """triangle() (main.rs:730-739) then push_triangle."""
        _check(load_library().b200rt_world_push_flat_triangle(self.world._h, self.object_index, _f32(positions, 9),
                                                             _f32(uvs, 6)), "push_flat_triangle")
        return self

    def push_square(self, positions, uvs) -> "ObjectProxy":
        """push_triangles(&square(..)) (main.rs:741-746)."""
        _check(load_library().b200rt_world_push_square(self.world._h, self.object_index, _f32(positions, 12),
                                                      _f32(uvs, 8)), "push_square")
        return self

    def push_sphere(self, center, radius: float) -> "ObjectProxy":
        _check(load_library().b200rt_world_push_sphere(self.world._h, self.object_index, _f32(center, 3),
                                                      float(radius)), "push_sphere")
        return self

    def load_obj(self, path: str, scale_div: float = 3.0, offset=(0.7, 1.0, -0.5)) -> int:
        """push_triangles(&load_obj(path)) (main.rs:778-807, 810, 825). Returns the triangle count."""
        return _check(load_library().b200rt_world_load_obj(self.world._h, self.object_index, os.fsencode(path),
                                                          float(scale_div), _f32(offset, 3)), "load_obj")

    def load_obj_ex(self, path: str, scale_div: float = 3.0, offset=(0.7, 1.0, -0.5), model_index: int = 0,
                    use_texcoords: bool = False, use_normals: bool = False) -> int:
        """What tobj::load_obj returns beyond models[0].positions (main.rs:786-790): model `model_index` (-1 = all models),
        the file's `vt` as uv and `vn` as vertex normals.  Returns the triangle count."""
        flags = (OBJ_USE_TEXCOORDS if use_texcoords else 0) | (OBJ_USE_NORMALS if use_normals else 0)
        return _check(load_library().b200rt_world_load_obj_ex(self.world._h, self.object_index, os.fsencode(path),
                                                             float(scale_div), _f32(offset, 3), int(model_index), flags), "load_obj_ex")


class World:
    """World (main.rs:130-178): objects, triangles, spheres, lights in global push order."""

    def __init__(self):
        self._lib = load_library()
        self._h = C.c_void_p(self._lib.b200rt_world_new())
        if not self._h:
            raise MemoryError("b200rt_world_new failed")

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.b200rt_world_free(self._h)
                self._h = None
        except Exception:
            pass

    def push_object(self, material: Material) -> ObjectProxy:
        idx = _check(self._lib.b200rt_world_push_object(self._h, C.byref(material)), "push_object")
        return ObjectProxy(self, idx)

    def push_light(self, light: Light) -> None:
        _check(self._lib.b200rt_world_push_light(self._h, C.byref(light)), "push_light")

    def scene(self) -> Scene:
        """A view of the arrays (valid until the next push)."""
        s = Scene()
        _check(self._lib.b200rt_world_scene(self._h, C.byref(s)), "world_scene")
        return s

    @classmethod
    def fixture(cls, obj_path: Optional[str] = None) -> "World":
        """The scene literal of main() (main.rs:810-1075)."""
        w = cls()
        _check(w._lib.b200rt_world_fixture(w._h, os.fsencode(obj_path) if obj_path else None), "world_fixture")
        return w


def fixture_camera() -> Camera:
    cam = Camera()
    load_library().b200rt_fixture_camera(C.byref(cam))
    return cam


def default_params(**overrides) -> Params:
    p = Params()
    load_library().b200rt_default_params(C.byref(p))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def copy_params(p: Params, **overrides) -> Params:
    q = Params()
    C.memmove(C.byref(q), C.byref(p), C.sizeof(Params))
    for k, v in overrides.items():
        setattr(q, k, v)
    return q


# ---- GPU context ---------------------------------------------------------------------------------------
class Context:
    """One CUDA render context (one per GPU / process)."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        h = C.c_void_p()
        _check(self._lib.b200rt_create(int(device), C.byref(h)), "b200rt_create")
        self._h = h
        self.device = int(device)
        self._world = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200rt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_info(self):
        sm, khz, mem = C.c_int(), C.c_int(), C.c_size_t()
        _check(self._lib.b200rt_device_info(self._h, C.byref(sm), C.byref(khz), C.byref(mem)), "device_info")
        return {"sm_count": sm.value, "sm_clock_khz": khz.value, "hbm_bytes": mem.value}

    def upload_scene(self, world_or_scene) -> None:
        scene = world_or_scene.scene() if isinstance(world_or_scene, World) else world_or_scene
        _check(self._lib.b200rt_upload_scene(self._h, C.byref(scene)), "upload_scene", self)
        self._world = world_or_scene  # keep the host arrays alive

    # host-buffer entry points -------------------------------------------------------------------------
    def render_whitted(self, cam: Camera, params: Params, out_rgb: Optional[np.ndarray] = None,
                       want_prim_id: bool = True):
        h, w = params.height, params.width
        if out_rgb is None:
            out_rgb = np.zeros((h, w, 3), dtype=np.float32)
        assert out_rgb.dtype == np.float32 and out_rgb.shape == (h, w, 3) and out_rgb.flags.c_contiguous
        prim = np.full((h, w), -2, dtype=np.int32) if want_prim_id else None
        _check(self._lib.b200rt_render_whitted(self._h, C.byref(cam), C.byref(params), out_rgb.ctypes.data,
                                               prim.ctypes.data if prim is not None else None),
               "render_whitted", self)
        return out_rgb, prim

    def render_distributed(self, cam: Camera, params: Params, epoch_begin: int, epoch_count: int,
                           accum: Optional[np.ndarray] = None) -> np.ndarray:
        h, w = params.height, params.width
        if accum is None:
            accum = np.zeros((h, w, 4), dtype=np.float32)
        assert accum.dtype == np.float32 and accum.shape == (h, w, 4) and accum.flags.c_contiguous
        _check(self._lib.b200rt_render_distributed(self._h, C.byref(cam), C.byref(params), epoch_begin, epoch_count,
                                                   accum.ctypes.data), "render_distributed", self)
        return accum

    def intersect(self, rays: np.ndarray, cast_mode: int = CAST_TWO_PHASE) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        _check(self._lib.b200rt_intersect(self._h, rays.ctypes.data, rays.shape[0], cast_mode, hits.ctypes.data),
               "intersect", self)
        return hits

    # device-pointer entry points (torch tensors own the memory; torch is plumbing only) -----------------
    def render_whitted_device(self, cam: Camera, params: Params, d_rgb: int, d_prim: int = 0, stream: int = 0):
        _check(self._lib.b200rt_render_whitted_device(self._h, C.byref(cam), C.byref(params), d_rgb, d_prim or None,
                                                      stream or None), "render_whitted_device", self)

    def render_distributed_device(self, cam: Camera, params: Params, epoch_begin: int, epoch_count: int,
                                  d_accum: int, stream: int = 0):
        _check(self._lib.b200rt_render_distributed_device(self._h, C.byref(cam), C.byref(params), epoch_begin,
                                                          epoch_count, d_accum, stream or None),
               "render_distributed_device", self)

    def resolve_device(self, d_accum: int, d_rgb: int, n_pixels: int, stream: int = 0):
        _check(self._lib.b200rt_resolve_device(self._h, d_accum, d_rgb, n_pixels, stream or None), "resolve_device", self)

    def intersect_device(self, d_rays: int, n: int, d_hits: int, cast_mode: int = CAST_TWO_PHASE, stream: int = 0):
        _check(self._lib.b200rt_intersect_device(self._h, d_rays, n, cast_mode, d_hits, stream or None),
               "intersect_device", self)

    def post_process(self, rgb: np.ndarray):
        """main.rs:748-762 on the device (host buffers): returns (normalised copy, p98 divisor or 0)."""
        out = np.ascontiguousarray(rgb, dtype=np.float32).copy()
        p = C.c_float(0.0)
        _check(self._lib.b200rt_post_process(self._h, out.ctypes.data, out.size // 3, C.byref(p)), "post_process", self)
        return out, float(p.value)

    def post_process_device(self, d_rgb: int, n_pixels: int, d_p98: int = 0, stream: int = 0):
        _check(self._lib.b200rt_post_process_device(self._h, d_rgb, n_pixels, d_p98, stream), "post_process_device", self)

    def encode_srgb8(self, rgb: np.ndarray) -> np.ndarray:
        """image.rs:55-66 on the device (host buffers)."""
        src = np.ascontiguousarray(rgb, dtype=np.float32)
        out = np.empty(src.shape, dtype=np.uint8)
        _check(self._lib.b200rt_encode_srgb8(self._h, src.ctypes.data, src.size, out.ctypes.data), "encode_srgb8", self)
        return out

    def encode_srgb8_device(self, d_rgb: int, n_values: int, d_out: int, stream: int = 0):
        _check(self._lib.b200rt_encode_srgb8_device(self._h, d_rgb, n_values, d_out, stream), "encode_srgb8_device", self)

    def stats(self) -> dict:
        buf = (C.c_char * 256)()        # (room for the larger stats struct of an older tuning build, B200RT_LIB)
        s = Stats.from_buffer(buf)
        _check(self._lib.b200rt_get_stats(self._h, C.byref(s)), "get_stats", self)
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def set_kernel_timing(self, enabled: bool) -> None:
        _check(self._lib.b200rt_set_kernel_timing(self._h, 1 if enabled else 0), "set_kernel_timing", self)

    def reset_stats(self) -> None:
        _check(self._lib.b200rt_reset_stats(self._h), "reset_stats", self)

    def filter_bench(self, variant: int, blocks_per_sm: int = 4, iters: int = 256):
        """K2 micro-benchmark of the filter loop; returns (kernel_ms, pair_tests)."""
        ms, pairs = C.c_float(), C.c_uint64()
        _check(self._lib.b200rt_filter_bench(self._h, variant, blocks_per_sm, iters, C.byref(ms), C.byref(pairs)),
               "filter_bench", self)
        return ms.value, pairs.value

    def pipe_bench(self, variant: int):
        ms, ipc = C.c_float(), C.c_double()
        _check(self._lib.b200rt_pipe_bench(self._h, variant, C.byref(ms), C.byref(ipc)), "pipe_bench", self)
        return ms.value, ipc.value

    def measure_fp32_peak(self):
        t, mhz = C.c_double(), C.c_double()
        _check(self._lib.b200rt_measure_fp32_peak(self._h, C.byref(t), C.byref(mhz)), "measure_fp32_peak", self)
        return t.value, mhz.value


class Group:
    """A device group (include/b200rt.h, "device groups"): one frame sharded over several GPUs, NCCL inside the library.

    ``Group(devices=[0, 1, ...])`` drives every listed GPU from this process (ncclCommInitAll);
    ``Group.rank(device, rank, n_ranks, unique_id)`` is one rank of a one-process-per-GPU job — rank 0 makes the id with
    ``Group.unique_id()`` and the host program (torch.distributed, MPI, a file) hands it to the others.
    Render calls are collective; the frame lands on rank 0."""

    def __init__(self, devices: Optional[Sequence[int]] = None, _handle=None, _rank0: bool = True):
        self._lib = load_library()
        self._world = None
        if _handle is not None:
            self._h, self.is_root = _handle, _rank0
            return
        devs = list(devices if devices is not None else [0])
        arr = (C.c_int * len(devs))(*devs)
        h = C.c_void_p()
        _check(self._lib.b200rt_group_create(arr, len(devs), C.byref(h)), "b200rt_group_create")
        self._h, self.is_root = h, True

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_char * GROUP_ID_BYTES)()
        _check(load_library().b200rt_group_unique_id(buf, GROUP_ID_BYTES), "b200rt_group_unique_id")
        return bytes(buf)

    @classmethod
    def rank(cls, device: int, rank: int, n_ranks: int, unique_id: Optional[bytes]) -> "Group":
        lib = load_library()
        h = C.c_void_p()
        idb = (C.c_char * GROUP_ID_BYTES)(*(unique_id or b"\0" * GROUP_ID_BYTES)[:GROUP_ID_BYTES])
        _check(lib.b200rt_group_create_rank(int(device), int(rank), int(n_ranks), idb, GROUP_ID_BYTES, C.byref(h)),
               "b200rt_group_create_rank")
        return cls(_handle=h, _rank0=(rank == 0))

    def _ck(self, code: int, what: str) -> int:
        if code < 0:
            raise B200rtError(code, what, strerror(code) + ": " + self._lib.b200rt_group_last_error(self._h).decode())
        return code

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200rt_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self):
        n, l = C.c_int(), C.c_int()
        self._ck(self._lib.b200rt_group_size(self._h, C.byref(n), C.byref(l)), "group_size")
        return n.value, l.value

    def member(self, local_index: int = 0) -> "Context":
        """The single-GPU context of a local member (borrowed: stats, device info)."""
        h = C.c_void_p()
        self._ck(self._lib.b200rt_group_ctx(self._h, local_index, C.byref(h)), "group_ctx")
        ctx = Context.__new__(Context)
        ctx._lib, ctx._h, ctx.device, ctx._world = self._lib, h, -1, None
        ctx.close = lambda: None          # owned by the group
        return ctx

    def upload_scene(self, world_or_scene) -> None:
        scene = world_or_scene.scene() if isinstance(world_or_scene, World) else world_or_scene
        self._ck(self._lib.b200rt_group_upload_scene(self._h, C.byref(scene)), "group_upload_scene")
        self._world = world_or_scene

    def render_distributed(self, cam: Camera, params: Params, epoch_begin: int, epoch_count: int,
                           out_accum: Optional[np.ndarray] = None, by_rows: bool = False):
        """Epochs sharded over the ranks and summed on rank 0 (by_rows: rows sharded and gathered instead)."""
        h, w = params.height, params.width
        if out_accum is None and self.is_root:
            out_accum = np.zeros((h, w, 4), dtype=np.float32)
        ptr = out_accum.ctypes.data if out_accum is not None else None
        fn = self._lib.b200rt_group_render_distributed_rows if by_rows else self._lib.b200rt_group_render_distributed
        self._ck(fn(self._h, C.byref(cam), C.byref(params), epoch_begin, epoch_count, ptr), "group_render_distributed")
        return out_accum

    def render_distributed_device(self, cam: Camera, params: Params, epoch_begin: int, epoch_count: int, d_accum_root: int,
                                  by_rows: bool = False):
        fn = self._lib.b200rt_group_render_distributed_rows_device if by_rows else self._lib.b200rt_group_render_distributed_device
        self._ck(fn(self._h, C.byref(cam), C.byref(params), epoch_begin, epoch_count, C.c_void_p(d_accum_root or None)),
                 "group_render_distributed_device")

    def render_whitted(self, cam: Camera, params: Params, out_rgb: Optional[np.ndarray] = None, want_prim_id: bool = True):
        h, w = params.height, params.width
        prim = None
        if self.is_root:
            if out_rgb is None:
                out_rgb = np.zeros((h, w, 3), dtype=np.float32)
            prim = np.full((h, w), -2, dtype=np.int32) if want_prim_id else None
        self._ck(self._lib.b200rt_group_render_whitted(self._h, C.byref(cam), C.byref(params),
                                                       out_rgb.ctypes.data if out_rgb is not None else None,
                                                       prim.ctypes.data if prim is not None else None, 1 if want_prim_id else 0),
                 "group_render_whitted")
        return out_rgb, prim

    def render_whitted_device(self, cam: Camera, params: Params, d_rgb_root: int):
        self._ck(self._lib.b200rt_group_render_whitted_device(self._h, C.byref(cam), C.byref(params), C.c_void_p(d_rgb_root or None)),
                 "group_render_whitted_device")

    def last_render_ms(self) -> float:
        ms = C.c_float()
        self._ck(self._lib.b200rt_group_last_render_ms(self._h, C.byref(ms)), "group_last_render_ms")
        return ms.value


def render_main(ctx: "Context", cam: Camera, params: Params, epochs: int, out_path: Optional[str] = None,
                on_frame: Optional[Callable[[int, np.ndarray], None]] = None) -> np.ndarray:
    """The render part of the reference's `main()` (main.rs:1086-1173) on top of the C ABI, every stage on the GPU.

      1. the deterministic Whitted frame is added into `img` (main.rs:1089-1109), `post_process`ed in place (main.rs:1113)
         and written (main.rs:1114);
      2. every epoch adds its accepted samples (main.rs:1157-1167: a sample with a zero / subnormal / NaN / inf channel
         is dropped — the {sum, count} accumulator of one epoch holds exactly the accepted sample) into the SAME,
         already normalised image, which is normalised again by its p99 luma and written again (main.rs:1171-1172):
         the reference's progressive renormalisation chain ("accumulation quirk").

    `on_frame(k, img_u8)` is called after each write (k = 0 for the Whitted frame, 1.. for the epochs).  Returns the final
    linear image.  The epochs use the shared Philox sample stream, not rand 0.5's ISAAC."""
    h, w = params.height, params.width
    img = np.zeros((h, w, 3), dtype=np.float32)

    def finish(k: int) -> None:
        nonlocal img
        img, _ = ctx.post_process(img)                                   # main.rs:748-762
        if out_path is not None or on_frame is not None:
            u8 = ctx.encode_srgb8(img)                                   # image.rs:55-66
            if out_path is not None:
                write_png(out_path, u8)                                  # main.rs:764-776
            if on_frame is not None:
                on_frame(k, u8)

    rgb, _ = ctx.render_whitted(cam, params, want_prim_id=False)
    img = img + rgb                                                      # main.rs:1107
    finish(0)
    for i in range(epochs):                                              # main.rs:1129
        acc = ctx.render_distributed(cam, params, i, 1)
        img = img + acc[..., :3]                                         # main.rs:1165
        finish(i + 1)
    return img
