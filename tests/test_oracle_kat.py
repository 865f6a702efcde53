"""Hand-derived known-answer tests for the CPU oracle: the reference has no tests, so the
formulas themselves are the pins."""
import ctypes as C
import math

import numpy as np
import pytest


def f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def one_ray(b, o, d, face=0, ex=-1, exf=0):
    r = np.zeros(1, dtype=b.RAY_DTYPE)
    r["origin"] = o; r["direction"] = d; r["face_direction"] = face; r["exclude_prim"] = ex; r["exclude_face"] = exf
    return r


def simple_world(b):
    """one triangle in z=0 (normal +z), one unit sphere at (0,0,-5), material 0"""
    w = b.World()
    o = w.push_object(b.color_material())
    o.push_flat_triangle([[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 0], [1, 0], [0, 1]])
    o.push_sphere([0, 0, -5], 1.0)
    return w


def test_ray_through_sphere_centre(b200rt, oracle):
    w = simple_world(b200rt)
    # front hit: t = |c-o| - r ; back-face ray: t = |c-o| + r  (main.rs:273-275)
    h = oracle.intersect(w.scene(), one_ray(b200rt, [5, 5, 5], [0, 0, -1]))     # misses everything
    assert h["prim_id"][0] == -1
    h = oracle.intersect(w.scene(), one_ray(b200rt, [0.5, 2.0, 3], [0, 0, -1]))
    assert h["prim_id"][0] == -1
    h = oracle.intersect(w.scene(), one_ray(b200rt, [0, 0, -2], [0, 0, -1], face=b200rt.FACE_FRONT))
    assert h["prim_id"][0] == 1 and h["face_direction"][0] == b200rt.FACE_FRONT
    assert h["distance"][0] == pytest.approx(2.0)                 # |c-o| - r = 3 - 1
    np.testing.assert_allclose(h["normal"][0], [0, 0, 1], atol=1e-7)
    h = oracle.intersect(w.scene(), one_ray(b200rt, [0, 0, -2], [0, 0, -1], face=b200rt.FACE_BACK))
    assert h["prim_id"][0] == 1 and h["face_direction"][0] == b200rt.FACE_BACK
    assert h["distance"][0] == pytest.approx(4.0)                 # |c-o| + r
    np.testing.assert_allclose(h["normal"][0], [0, 0, 1], atol=1e-7)   # flipped normal of the far side (main.rs:307)
    # FACE_BOTH from inside the sphere: tc < k -> back face (main.rs:276-277)
    h = oracle.intersect(w.scene(), one_ray(b200rt, [0, 0, -5], [1, 0, 0], face=b200rt.FACE_BOTH))
    assert h["face_direction"][0] == b200rt.FACE_BACK and h["distance"][0] == pytest.approx(1.0)
    # sphere uv (main.rs:310-313): normal (0,0,1) -> (acos(0)/pi, atan2(1,0)/(2pi)+0.5) = (0.5, 0.75)
    h = oracle.intersect(w.scene(), one_ray(b200rt, [0, 0, -2], [0, 0, -1]))
    np.testing.assert_allclose(h["uv"][0], [0.5, 0.75], atol=1e-6)


def test_unit_triangle_barycentrics_and_culling(b200rt, oracle):
    w = simple_world(b200rt)
    # front face looks toward +z; a ray travelling -z hits it from the front
    h = oracle.intersect(w.scene(), one_ray(b200rt, [0.25, 0.5, 2.0], [0, 0, -1]))
    assert h["prim_id"][0] == 0 and h["face_direction"][0] == b200rt.FACE_FRONT
    assert h["distance"][0] == pytest.approx(2.0)
    np.testing.assert_allclose(h["position"][0], [0.25, 0.5, 0.0], atol=1e-7)
    # uv = b0*uv0 + b1*uv1 + b2*uv2 with uv = vertex xy  =>  uv == hit xy (main.rs:235-252)
    np.testing.assert_allclose(h["uv"][0], [0.25, 0.5], atol=1e-6)
    np.testing.assert_allclose(h["normal"][0], [0, 0, 1], atol=1e-6)
    # from behind: Front rays cull it (main.rs:185), Back rays hit it with a negated normal (main.rs:250)
    assert oracle.intersect(w.scene(), one_ray(b200rt, [0.25, 0.5, -2.0], [0, 0, 1]))["prim_id"][0] == -1
    hb = oracle.intersect(w.scene(), one_ray(b200rt, [0.25, 0.5, -2.0], [0, 0, 1], face=b200rt.FACE_BACK))
    assert hb["prim_id"][0] == 0 and hb["face_direction"][0] == b200rt.FACE_BACK
    np.testing.assert_allclose(hb["normal"][0], [0, 0, -1], atol=1e-6)
    # edges and vertices are inside (area >= 0 accepted, main.rs:224); just outside is not
    assert oracle.intersect(w.scene(), one_ray(b200rt, [0.5, 0.0, 1.0], [0, 0, -1]))["prim_id"][0] == 0
    assert oracle.intersect(w.scene(), one_ray(b200rt, [0.0, 0.0, 1.0], [0, 0, -1]))["prim_id"][0] == 0
    assert oracle.intersect(w.scene(), one_ray(b200rt, [0.5, -1e-6, 1.0], [0, 0, -1]))["prim_id"][0] != 0  # falls through to the sphere
    # t <= 0 is rejected (main.rs:205): origin on the plane
    assert oracle.intersect(w.scene(), one_ray(b200rt, [0.25, 0.25, 0.0], [0, 0, -1]))["prim_id"][0] != 0


def test_exclusion_semantics(b200rt, oracle):
    w = simple_world(b200rt)
    o, d = [0.25, 0.5, 2.0], [0, 0, -1]
    # exclusion {tri 0, Front} removes the front-face hit, {tri 0, Back} does not (main.rs:190-200)
    h = oracle.intersect(w.scene(), one_ray(b200rt, o, d, ex=0, exf=b200rt.FACE_FRONT))
    assert h["prim_id"][0] == 1          # falls through to the sphere behind the triangle
    h = oracle.intersect(w.scene(), one_ray(b200rt, o, d, ex=0, exf=b200rt.FACE_BACK))
    assert h["prim_id"][0] == 0
    h = oracle.intersect(w.scene(), one_ray(b200rt, o, d, ex=0, exf=b200rt.FACE_BOTH))
    assert h["prim_id"][0] == 1
    h = oracle.intersect(w.scene(), one_ray(b200rt, o, d, ex=1, exf=b200rt.FACE_BOTH))   # other primitive
    assert h["prim_id"][0] == 0


def test_tie_break_later_primitive_wins(b200rt, oracle):
    """nearest uses `nearest_t < t => skip` (main.rs:229-233, 298-302): exact ties go to the later primitive,
    and spheres are tested after all triangles."""
    w = b200rt.World()
    o = w.push_object(b200rt.color_material())
    tri = [[-1, -1, 0], [1, -1, 0], [0, 1, 0]]
    o.push_flat_triangle(tri)
    o.push_flat_triangle(tri)          # coincident copy
    h = oracle.intersect(w.scene(), one_ray(b200rt, [0, 0, 1], [0, 0, -1]))
    assert h["prim_id"][0] == 1
    o.push_sphere([0, 0, -1], 1.0)     # touches z=0 at the origin: t = 1 as well
    h = oracle.intersect(w.scene(), one_ray(b200rt, [0, 0, 1], [0, 0, -1]))
    assert h["distance"][0] == 1.0 and h["prim_id"][0] == 2


def test_refract_closure(oracle):
    lib = oracle.load()
    out = (C.c_float * 3)()
    n = f3([0, 0, 1])
    l = np.array([0.6, 0.0, -0.8], dtype=np.float32)
    # k = 1: passthrough
    assert lib.oracle_refract(n, f3(l), 1.0, out) == 1
    np.testing.assert_allclose(out[:], l, atol=1e-6)
    # Snell: sin_t = sin_i / k
    assert lib.oracle_refract(n, f3(l), 1.5, out) == 1
    assert math.hypot(out[0], out[1]) == pytest.approx(0.6 / 1.5, abs=1e-6)
    assert out[2] < 0
    # total internal reflection threshold k^2 >= 1 - cos^2 (main.rs:346): sin_i = 0.6
    assert lib.oracle_refract(n, f3(l), 0.61, out) == 1
    assert lib.oracle_refract(n, f3(l), 0.59, out) == 0


def test_from_arc(oracle):
    lib = oracle.load()
    out = (C.c_float * 3)()
    z = f3([0, 0, 1])
    v = f3([0.3, -0.2, 0.9])
    lib.oracle_from_arc_rotate(z, z, v, out)                      # identity
    assert list(out) == list(v)
    lib.oracle_from_arc_rotate(z, f3([1, 0, 0]), z, out)          # z -> x
    np.testing.assert_allclose(out[:], [1, 0, 0], atol=1e-6)
    lib.oracle_from_arc_rotate(z, f3([0, 0, -1]), z, out)         # antiparallel: pi about norm(x^ x z) = -y
    np.testing.assert_allclose(out[:], [0, 0, -1], atol=1e-6)
    lib.oracle_from_arc_rotate(z, f3([0, 0, -1]), f3([1, 0, 0]), out)
    np.testing.assert_allclose(out[:], [-1, 0, 0], atol=1e-6)


def test_lights(b200rt, oracle):
    lib = oracle.load()
    d, c, o = (C.c_float * 3)(), (C.c_float * 3)(), (C.c_float * 3)()
    ho = C.c_int()
    spot = b200rt.spot_light([0, 10, 0], [0, -1, 0], math.radians(60.0), 1.0, [1.0, 0.5, 0.9])
    # on the axis, 10 below: angular attenuation 1, distance attenuation 1/10 (NOT 1/d^2; lights.rs:64)
    assert lib.oracle_light_approx(C.byref(spot), f3([0, 0, 0]), d, c, o, C.byref(ho)) == 1
    np.testing.assert_allclose(d[:], [0, -1, 0], atol=1e-7)
    np.testing.assert_allclose(c[:], [0.1, 0.05, 0.09], rtol=1e-5)
    assert ho.value == 1
    # half-way out (30 deg): (1 - 0.5)^(1+eps) / dist
    p = [10 * math.tan(math.radians(30.0)), 0, 0]
    assert lib.oracle_light_approx(C.byref(spot), f3(p), d, c, o, C.byref(ho)) == 1
    dist = math.hypot(p[0], 10)
    np.testing.assert_allclose(c[:], np.array([1.0, 0.5, 0.9]) * 0.5 / dist, rtol=1e-4)
    # outside the cone -> None (lights.rs:59-61)
    assert lib.oracle_light_approx(C.byref(spot), f3([10 * math.tan(math.radians(61.0)), 0, 0]), d, c, o, C.byref(ho)) == 0
    point = b200rt.point_light([0, 0.1, 0], [0.8, 0.8, 1.0])
    assert lib.oracle_light_approx(C.byref(point), f3([0, 2.1, 0]), d, c, o, C.byref(ho)) == 1
    np.testing.assert_allclose(c[:], [0.4, 0.4, 0.5], rtol=1e-6)
    np.testing.assert_allclose(d[:], [0, 1, 0], atol=1e-7)
    dl = b200rt.directional_light([0, -1, 0], [1, 1, 1])
    assert lib.oracle_light_approx(C.byref(dl), f3([3, 4, 5]), d, c, o, C.byref(ho)) == 1
    assert ho.value == 0 and list(c) == [1, 1, 1]


def test_materials(b200rt, oracle):
    lib = oracle.load()
    out = (C.c_float * 3)()
    m = b200rt.color_material(diffuse_color=(1.0, 0.8, 0.6), smoothness=0.5, specular_color=(1, 1, 0))
    n = f3([0, 0, 1])
    lib.oracle_get_diffuse(C.byref(m), n, f3([0, 0.6, 0.8]), out)
    np.testing.assert_allclose(out[:], [0.8, 0.64, 0.48], rtol=1e-6)
    lib.oracle_get_diffuse(C.byref(m), n, f3([0, 0.6, -0.8]), out)
    assert list(out) == [0, 0, 0]
    # mirror configuration: R.V = 1 -> amount = (s+8)/(8 pi), s = 1/(smoothness+eps)  (materials.rs:60-64)
    lib.oracle_get_specular(C.byref(m), n, f3([0, -0.6, 0.8]), f3([0, 0.6, 0.8]), out)
    s = 1.0 / (0.5 + 1.1920929e-7)
    np.testing.assert_allclose(out[:], np.array([1, 1, 0]) * (s + 8) / (8 * math.pi), rtol=1e-5)
    # procedural closures (main.rs:848-863, 1019-1026)
    wall = b200rt.generative_material(b200rt.DIFFUSE_STRIPE_V, b200rt.NORMAL_SINCOS_U, freq=20.0, c0=(1, 1, 1),
                                     c1=(0.5, 0.5, 1.0), nfreq=10.0)
    res = b200rt.Material()
    uv = (C.c_float * 2)(0.0125, 0.04)          # uv.y*20 = 0.8 -> 0 (even) ; angle = 0.0125*20*pi = pi/4
    lib.oracle_material_approx(C.byref(wall), uv, C.byref(res))
    assert list(res.diffuse_color) == [1, 1, 1]
    np.testing.assert_allclose(list(res.normal), [math.sin(math.pi / 4), 0, math.cos(math.pi / 4)], atol=1e-6)
    uv = (C.c_float * 2)(0.05, 0.06)            # uv.y*20 = 1.2 -> 1 (odd) ; angle = pi -> (0,0,-1) flipped to +z
    lib.oracle_material_approx(C.byref(wall), uv, C.byref(res))
    np.testing.assert_allclose(list(res.diffuse_color), [0.5, 0.5, 1.0])
    assert res.normal[2] > 0.999
    uv = (C.c_float * 2)(0.0, -0.06)            # negative: `as i32` truncates toward zero, % keeps sign: -1 % 2 = -1 != 0
    lib.oracle_material_approx(C.byref(wall), uv, C.byref(res))
    np.testing.assert_allclose(list(res.diffuse_color), [0.5, 0.5, 1.0])
    checker = b200rt.generative_material(b200rt.DIFFUSE_CHECKER_UPV, freq=10.0, c0=(1, 0.1, 0.1), c1=(0.1, 0.1, 1))
    uv = (C.c_float * 2)(0.1, 0.15)             # (0.25)*10 = 2.5 -> 2 even
    lib.oracle_material_approx(C.byref(checker), uv, C.byref(res))
    np.testing.assert_allclose(list(res.diffuse_color), [1, 0.1, 0.1], rtol=1e-6)


def test_philox_known_answers(oracle):
    lib = oracle.load()
    out = (C.c_uint32 * 4)()
    lib.oracle_philox4x32_10((C.c_uint32 * 4)(0, 0, 0, 0), (C.c_uint32 * 2)(0, 0), out)
    assert [hex(v) for v in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    lib.oracle_philox4x32_10((C.c_uint32 * 4)(*[0xffffffff] * 4), (C.c_uint32 * 2)(0xffffffff, 0xffffffff), out)
    assert [hex(v) for v in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    lib.oracle_philox4x32_10((C.c_uint32 * 4)(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344),
                             (C.c_uint32 * 2)(0xa4093822, 0x299f31d0), out)
    assert [hex(v) for v in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_sample_stream(oracle):
    lib = oracle.load()
    a = (C.c_float * 9)(); bb = (C.c_float * 9)()
    lib.oracle_sample_uniforms(5, 10, 20, 3, 9, a)
    lib.oracle_sample_uniforms(5, 10, 20, 3, 9, bb)
    assert list(a) == list(bb)
    assert all(0.0 <= v < 1.0 for v in a)
    lib.oracle_sample_uniforms(5, 10, 20, 4, 9, bb)     # another epoch: independent stream
    assert list(a) != list(bb)
    lib.oracle_sample_uniforms(5, 20, 10, 3, 9, bb)     # (y,x) swapped
    assert list(a) != list(bb)
    big = (C.c_float * 4096)()
    lib.oracle_sample_uniforms(1, 0, 0, 0, 4096, big)
    assert abs(np.mean(big[:]) - 0.5) < 0.02


def test_camera_shoot(b200rt, oracle):
    """main.rs:84-99 with the fixture camera: centre pixel looks along `toward`, origin is BEHIND the centre
    (near = -0.1)."""
    cam = b200rt.fixture_camera()
    r = b200rt.Ray()
    oracle.load().oracle_camera_shoot(C.byref(cam), 0.0, 0.0, C.byref(r))
    t = -1 / math.sqrt(3)
    np.testing.assert_allclose(list(r.direction), [t, t, t], atol=1e-6)
    np.testing.assert_allclose(list(r.origin), [2 + 0.1 / math.sqrt(3), 2.5 + 0.1 / math.sqrt(3), 2 + 0.1 / math.sqrt(3)], atol=1e-6)
    oracle.load().oracle_camera_shoot(C.byref(cam), 0.0, 0.5, C.byref(r))
    # top of the frame: angle to the axis = atan(0.5 * tan(30 deg)) = 16.1 deg (the "vFOV 32.2" quirk)
    cosang = sum(a * t for a in r.direction)
    assert math.degrees(math.acos(cosang)) == pytest.approx(math.degrees(math.atan(0.5 * math.tan(math.radians(30)))), abs=1e-3)
    assert r.direction[1] > t


def test_post_process_srgb_and_accumulator(oracle):
    img = np.zeros((100, 3), dtype=np.float32)
    img[:, :] = np.linspace(0.01, 1.0, 100, dtype=np.float32)[:, None]
    out, p99 = oracle.post_process(img)
    assert p99 == pytest.approx(1.0, abs=1e-6)          # luma coefficients sum to 1; index int(100*0.99) = 99
    img[50] = 0.0                                        # a zero pixel is not `is_normal`: excluded from the sort
    out, p99 = oracle.post_process(img)
    assert p99 == pytest.approx(1.0, abs=1e-6)
    enc = oracle.encode_srgb8(np.array([0.0, 0.0031308, 0.5, 1.0, 2.0, -1.0], dtype=np.float32))
    assert enc.tolist() == [0, 10, 188, 255, 255, 0]
    acc = np.array([[3.0, 6.0, 9.0, 3.0], [1.0, 1.0, 1.0, 0.0]], dtype=np.float32)   # photon.rs:18-21
    np.testing.assert_allclose(oracle.resolve(acc), [[1, 2, 3], [0, 0, 0]])


def test_is_normal_filter_drops_black_samples(b200rt, oracle, fixture_world):
    """main.rs:1157-1160: a sample with ANY zero channel is dropped, so background pixels never count."""
    cam = b200rt.fixture_camera()
    params = b200rt.default_params(width=64, height=48, seed=1)
    acc, cnt = oracle.render_distributed(fixture_world.scene(), cam, params, 0, 3)
    assert acc[..., 3].max() <= 3
    assert cnt["samples"] == int(acc[..., 3].sum())
    black = acc[..., 3] == 0
    assert black.any() and np.all(acc[black][:, :3] == 0)
    s = oracle.sample_distributed(fixture_world.scene(), cam, params, 24, 32, 0)
    assert np.isfinite(s).all()
