"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE (the checker).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")

if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import __graft_entry__ as _ge  # noqa: E402

b = _ge.load_package()

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    src_newer = (not os.path.exists(ORACLE_LIB)) or any(
        os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(ORACLE_LIB) for f in ("oracle.cpp", "oracle.h"))
    if src_newer:
        res = subprocess.run(["make", "-C", ORACLE_DIR], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    lib = C.CDLL(ORACLE_LIB)
    vp = C.c_void_p
    u64p = C.POINTER(C.c_uint64)
    f32p = C.POINTER(C.c_float)
    lib.oracle_render_whitted.restype = C.c_int
    lib.oracle_render_whitted.argtypes = [C.POINTER(b.Scene), C.POINTER(b.Camera), C.POINTER(b.Params), vp, vp, u64p, C.c_int]
    lib.oracle_render_distributed.restype = C.c_int
    lib.oracle_render_distributed.argtypes = [C.POINTER(b.Scene), C.POINTER(b.Camera), C.POINTER(b.Params), C.c_uint32,
                                              C.c_uint32, vp, u64p, C.c_int]
    lib.oracle_intersect.restype = C.c_int
    lib.oracle_intersect.argtypes = [C.POINTER(b.Scene), vp, C.c_size_t, vp]
    lib.oracle_sample_distributed.restype = C.c_int
    lib.oracle_sample_distributed.argtypes = [C.POINTER(b.Scene), C.POINTER(b.Camera), C.POINTER(b.Params), C.c_uint32,
                                              C.c_uint32, C.c_uint32, f32p]
    lib.oracle_post_process.restype = C.c_float
    lib.oracle_post_process.argtypes = [vp, C.c_size_t]
    lib.oracle_encode_srgb8.restype = None
    lib.oracle_encode_srgb8.argtypes = [vp, C.c_size_t, vp]
    lib.oracle_resolve.restype = None
    lib.oracle_resolve.argtypes = [vp, C.c_size_t, vp]
    lib.oracle_camera_shoot.restype = None
    lib.oracle_camera_shoot.argtypes = [C.POINTER(b.Camera), C.c_float, C.c_float, C.POINTER(b.Ray)]
    lib.oracle_refract.restype = C.c_int
    lib.oracle_refract.argtypes = [f32p, f32p, C.c_float, f32p]
    lib.oracle_from_arc_rotate.restype = None
    lib.oracle_from_arc_rotate.argtypes = [f32p, f32p, f32p, f32p]
    lib.oracle_light_approx.restype = C.c_int
    lib.oracle_light_approx.argtypes = [C.POINTER(b.Light), f32p, f32p, f32p, f32p, C.POINTER(C.c_int)]
    lib.oracle_material_approx.restype = None
    lib.oracle_material_approx.argtypes = [C.POINTER(b.Material), f32p, C.POINTER(b.Material)]
    lib.oracle_get_diffuse.restype = None
    lib.oracle_get_diffuse.argtypes = [C.POINTER(b.Material), f32p, f32p, f32p]
    lib.oracle_get_specular.restype = None
    lib.oracle_get_specular.argtypes = [C.POINTER(b.Material), f32p, f32p, f32p, f32p]
    lib.oracle_philox4x32_10.restype = None
    lib.oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.oracle_sample_uniforms.restype = None
    lib.oracle_sample_uniforms.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, f32p]
    lib.oracle_max_threads.restype = C.c_int
    lib.oracle_max_threads.argtypes = []
    _lib = lib
    return lib


class GoldenFixture:
    """The reference's scene literal, camera and default parameters from tests/golden/fixture_scene.npz: plain ctypes
    records backed by numpy arrays — libb200rt.so is not loaded (bench.py's reference arm runs the oracle only)."""

    def __init__(self):
        z = np.load(os.path.join(ROOT, "tests", "golden", "fixture_scene.npz"))
        self._keep = {k: np.ascontiguousarray(z[k]) for k in ("triangles", "spheres", "materials", "lights")}
        s = b.Scene()
        for name, typ, ptr_field, n_field in (("triangles", b.Triangle, "triangles", "n_triangles"),
                                              ("spheres", b.Sphere, "spheres", "n_spheres"),
                                              ("materials", b.Material, "materials", "n_materials"),
                                              ("lights", b.Light, "lights", "n_lights")):
            arr = self._keep[name]
            setattr(s, ptr_field, C.cast(arr.ctypes.data, C.POINTER(typ)))
            setattr(s, n_field, arr.size // C.sizeof(typ))
        self.scene = s
        self.camera = b.Camera.from_buffer_copy(z["camera"].tobytes())
        self._params = z["params"].tobytes()

    def params(self, **overrides):
        p = b.Params.from_buffer_copy(self._params)
        for k, v in overrides.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
        return p


def host_threads() -> int:
    """Host cores this process may run on (torchrun exports OMP_NUM_THREADS=1: not what omp_get_max_threads says)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def max_threads() -> int:
    return load().oracle_max_threads()


def render_whitted(scene, cam, params, n_threads: int = 0):
    lib = load()
    h, w = params.height, params.width
    rgb = np.zeros((h, w, 3), dtype=np.float32)
    prim = np.full((h, w), -2, dtype=np.int32)
    cnt = (C.c_uint64 * 4)()
    rc = lib.oracle_render_whitted(C.byref(scene), C.byref(cam), C.byref(params), rgb.ctypes.data, prim.ctypes.data, cnt,
                                   n_threads)
    assert rc == 0, rc
    return rgb, prim, {"casts": cnt[0], "tri_pairs": cnt[1], "sph_pairs": cnt[2], "samples": cnt[3]}


def render_distributed(scene, cam, params, epoch_begin, epoch_count, accum=None, n_threads: int = 0):
    lib = load()
    h, w = params.height, params.width
    if accum is None:
        accum = np.zeros((h, w, 4), dtype=np.float32)
    cnt = (C.c_uint64 * 4)()
    rc = lib.oracle_render_distributed(C.byref(scene), C.byref(cam), C.byref(params), epoch_begin, epoch_count,
                                       accum.ctypes.data, cnt, n_threads)
    assert rc == 0, rc
    return accum, {"casts": cnt[0], "tri_pairs": cnt[1], "sph_pairs": cnt[2], "samples": cnt[3]}


def intersect(scene, rays: np.ndarray) -> np.ndarray:
    lib = load()
    rays = np.ascontiguousarray(rays, dtype=b.RAY_DTYPE)
    hits = np.zeros(rays.shape[0], dtype=b.HIT_DTYPE)
    rc = lib.oracle_intersect(C.byref(scene), rays.ctypes.data, rays.shape[0], hits.ctypes.data)
    assert rc == 0, rc
    return hits


def sample_distributed(scene, cam, params, y, x, epoch):
    out = (C.c_float * 3)()
    rc = load().oracle_sample_distributed(C.byref(scene), C.byref(cam), C.byref(params), y, x, epoch, out)
    assert rc == 0
    return np.array(out[:], dtype=np.float32)


def post_process(rgb: np.ndarray):
    out = np.ascontiguousarray(rgb, dtype=np.float32).copy()
    p = load().oracle_post_process(out.ctypes.data, out.size // 3)
    return out, p


def encode_srgb8(rgb: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(rgb, dtype=np.float32)
    out = np.zeros(src.shape, dtype=np.uint8)
    load().oracle_encode_srgb8(src.ctypes.data, src.size, out.ctypes.data)
    return out


def resolve(accum: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(accum, dtype=np.float32)
    out = np.zeros(src.shape[:-1] + (3,), dtype=np.float32)
    load().oracle_resolve(src.ctypes.data, src.size // 4, out.ctypes.data)
    return out
