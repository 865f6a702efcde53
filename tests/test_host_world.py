"""Host-side scene construction (World / ObjectProxy / triangle() / square() / load_obj; main.rs:161-178,
705-746, 778-807) and the scene literal of main() (main.rs:810-1075)."""
import ctypes as C

import numpy as np
import pytest


def tri_array(b200rt, scene):
    raw = C.string_at(scene.triangles, C.sizeof(b200rt.Triangle) * scene.n_triangles)
    return np.frombuffer(raw, dtype=np.float32).reshape(scene.n_triangles, 25)


def test_fixture_scene_census(b200rt, fixture_world):
    s = fixture_world.scene()
    assert (s.n_triangles, s.n_spheres, s.n_materials, s.n_lights) == (64, 4, 9, 3)
    t = tri_array(b200rt, s)
    obj = t[:, 24].view(np.uint32)
    assert obj[:36].tolist() == [0] * 36                 # dodecahedron
    assert obj[36:38].tolist() == [1, 1]                 # floor
    assert obj[38:40].tolist() == [2, 2]                 # wall
    assert obj[40:52].tolist() == [3] * 12 and obj[52:64].tolist() == [4] * 12     # glass slabs
    assert [s.spheres[i].object_index for i in range(4)] == [5, 6, 7, 8]
    np.testing.assert_allclose(list(s.spheres[0].center), [-0.5, 0.5, 0.5 / np.sqrt(np.float32(3))], rtol=1e-7)
    np.testing.assert_allclose(list(s.spheres[3].center), [0.0, 0.5 + np.sqrt(np.float32(2) / np.float32(3)), 0.0], rtol=1e-7)
    m = s.materials
    assert m[2].kind == b200rt.MATERIAL_GENERATIVE and m[2].diffuse_fn == b200rt.DIFFUSE_STRIPE_V
    assert m[2].normal_fn == b200rt.NORMAL_SINCOS_U and m[2].fn_params[0] == 20.0 and m[2].fn_params[7] == 10.0
    assert m[7].diffuse_fn == b200rt.DIFFUSE_CHECKER_UPV and list(m[7].specular_color) == [0, 0, 1]
    assert m[3].refraction_index == pytest.approx(1.6) and m[3].transparency == 1.0 and m[6].transparency == pytest.approx(0.96)
    assert [s.lights[i].kind for i in range(3)] == [b200rt.LIGHT_DIRECTIONAL, b200rt.LIGHT_SPOT, b200rt.LIGHT_POINT]
    assert s.lights[0].has_origin == 0 and s.lights[1].angle == pytest.approx(np.pi / 3, rel=1e-6)
    # the floor quad repeats uv (0,1) on its 4th corner, as the reference does (main.rs:843)
    assert t[37, 22:24].tolist() == [0.0, 1.0]


def test_flat_normals_and_square_order(b200rt):
    w = b200rt.World()
    o = w.push_object(b200rt.color_material())
    o.push_square([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], [[0, 0], [1, 0], [1, 1], [0, 1]])
    t = tri_array(b200rt, w.scene())
    assert t.shape[0] == 2
    # square() = triangle(0,1,2), triangle(0,2,3)  (main.rs:741-746)
    assert t[0, [0, 1, 8, 9, 16, 17]].tolist() == [0, 0, 1, 0, 1, 1]
    assert t[1, [0, 1, 8, 9, 16, 17]].tolist() == [0, 0, 1, 1, 0, 1]
    # triangle(): normal = normalize((v1-v0) x (v2-v1)) on all three vertices (main.rs:731-738)
    for k in range(2):
        for v in range(3):
            assert t[k, 8 * v + 3: 8 * v + 6].tolist() == [0, 0, 1]
    assert t[1, 22:24].tolist() == [0, 1]


def test_push_order_defines_primitive_ids(b200rt, oracle):
    """PrimitiveIndex::Triangle(i) / Sphere(i) index GLOBAL push-order arrays (main.rs:706-720)."""
    w = b200rt.World()
    a = w.push_object(b200rt.color_material())
    b = w.push_object(b200rt.color_material())
    b.push_sphere([0, 0, -10], 1.0)
    a.push_flat_triangle([[-1, -1, -5], [1, -1, -5], [0, 1, -5]])
    b.push_flat_triangle([[-1, -1, -3], [1, -1, -3], [0, 1, -3]])
    s = w.scene()
    assert [s.triangles[i].object_index for i in range(2)] == [0, 1]
    rays = np.zeros(1, dtype=b200rt.RAY_DTYPE)
    rays["origin"] = [0, 0, 0]; rays["direction"] = [0, 0, -1]; rays["exclude_prim"] = -1
    h = oracle.intersect(s, rays)
    assert h["prim_id"][0] == 1 and h["object_index"][0] == 1 and h["distance"][0] == 3.0
    rays["exclude_prim"] = 1; rays["exclude_face"] = b200rt.FACE_BOTH
    assert oracle.intersect(s, rays)["prim_id"][0] == 0
    rays["origin"] = [5, 5, 0]
    assert oracle.intersect(s, rays)["prim_id"][0] == -1


OBJ_TEXT = """# comment
o first
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
vt 0 0
vn 0 0 1
f 1/1/1 2/1/1 3/1/1 4/1/1
f -4//1 -3//1 -2//1
g second
v 5 5 5
f 1 2 5
"""


def test_obj_import(b200rt, tmp_path):
    """tobj behaviours load_obj relies on (main.rs:784-790): first model only, fan triangulation, a/b/c forms,
    negative indices; plus the reference's fixed transform p/3 + (0.7, 1.0, -0.5) (main.rs:802) and uv = (0,0)."""
    path = tmp_path / "t.obj"
    path.write_text(OBJ_TEXT)
    w = b200rt.World()
    o = w.push_object(b200rt.color_material())
    n = o.load_obj(str(path))
    assert n == 3                                             # quad -> 2 triangles, + 1; the `g second` face is ignored
    t = tri_array(b200rt, w.scene())
    f32 = np.float32
    exp0 = [f32(0) / f32(3) + f32(0.7), f32(0) / f32(3) + f32(1.0), f32(0) / f32(3) + f32(-0.5)]
    assert t[0, 0:3].tolist() == [float(v) for v in exp0]
    exp1 = [f32(1) / f32(3) + f32(0.7), f32(1.0), f32(-0.5)]
    assert t[0, 8:11].tolist() == [float(v) for v in exp1]
    assert t[1, 16:19].tolist() == [float(f32(0.7)), float(f32(1) / f32(3) + f32(1.0)), float(f32(-0.5))]   # fan: (0,2,3)
    assert np.all(t[:, [6, 7, 14, 15, 22, 23]] == 0)          # uv = (0,0), main.rs:797-799
    assert np.array_equal(t[2, [0, 1, 2]], t[0, [0, 1, 2]])    # -4 resolves to vertex 1
    # no transform
    w2 = b200rt.World()
    o2 = w2.push_object(b200rt.color_material())
    assert o2.load_obj(str(path), scale_div=1.0, offset=(0, 0, 0)) == 3
    assert tri_array(b200rt, w2.scene())[0, 8:11].tolist() == [1, 0, 0]


OBJ_FULL = """# two models, texture coordinates and normals (what tobj::load_obj returns and load_obj drops)
o first
v 0 0 0
v 3 0 0
v 3 3 0
v 0 3 0
vt 0.25 0.5
vt 1 0
vt 1 1
vn 0 0 1
vn 0 1 0
f 1/1/1 2/2/1 3/3/2
f 1//1 3//1 4//1
g second
v 0 0 3
f 1/1 2/2 5/-1
f -1 -2 -3 -4
"""


def test_obj_import_models_texcoords_normals(b200rt, tmp_path):
    """N3: the tobj behaviours beyond models[0].positions - model selection, `vt`, `vn`, mixed corner forms."""
    path = tmp_path / "full.obj"
    path.write_text(OBJ_FULL)
    lib = b200rt.load_library()
    assert lib.b200rt_obj_model_count(str(path).encode()) == 2
    ident = dict(scale_div=1.0, offset=(0, 0, 0))
    # flags = 0, model 0 is load_obj
    wa, wb = b200rt.World(), b200rt.World()
    oa, ob = wa.push_object(b200rt.color_material()), wb.push_object(b200rt.color_material())
    assert oa.load_obj(str(path)) == 2 and ob.load_obj_ex(str(path)) == 2
    assert np.array_equal(tri_array(b200rt, wa.scene()), tri_array(b200rt, wb.scene()))
    # model 1 alone: its triangle and the quad (a fan of two) over the last four vertices
    w1 = b200rt.World()
    o1 = w1.push_object(b200rt.color_material())
    assert o1.load_obj_ex(str(path), model_index=1, **ident) == 3
    t = tri_array(b200rt, w1.scene())
    assert t[0, 16:19].tolist() == [0, 0, 3]                                  # corner 5 = the vertex pushed in `second`
    assert t[1, 0:3].tolist() == [0, 0, 3] and t[1, 8:11].tolist() == [0, 3, 0] and t[1, 16:19].tolist() == [3, 3, 0]   # -1 -2 -3
    assert t[2, 16:19].tolist() == [3, 0, 0]                                  # fan: (-1, -3, -4)
    # all models, with the file's uv and normals
    w2 = b200rt.World()
    o2 = w2.push_object(b200rt.color_material())
    assert o2.load_obj_ex(str(path), model_index=-1, use_texcoords=True, use_normals=True, **ident) == 5
    t = tri_array(b200rt, w2.scene())
    assert t[0, 6:8].tolist() == [0.25, 0.5] and t[0, 14:16].tolist() == [1, 0] and t[0, 22:24].tolist() == [1, 1]
    assert t[0, 3:6].tolist() == [0, 0, 1] and t[0, 19:22].tolist() == [0, 1, 0]      # per-corner vn
    assert np.all(t[1, [6, 7, 14, 15, 22, 23]] == 0) and t[1, 3:6].tolist() == [0, 0, 1]   # a//c: no uv, own normal
    assert t[2, 22:24].tolist() == [1, 1]                                     # vt -1 = the last vt
    flat = tri_array(b200rt, w1.scene())
    assert np.array_equal(t[2, 3:6], flat[0, 3:6])                            # no vn on a corner: triangle()'s flat normal
    # the reference's transform applies to every model alike
    w3 = b200rt.World()
    o3 = w3.push_object(b200rt.color_material())
    assert o3.load_obj_ex(str(path), model_index=1) == 3
    f32 = np.float32
    assert tri_array(b200rt, w3.scene())[0, 16:19].tolist() == [float(f32(0.7)), float(f32(1.0)), float(f32(3) / f32(3) + f32(-0.5))]
    with pytest.raises(b200rt.B200rtError):
        o3.load_obj_ex(str(path), model_index=2)
    assert lib.b200rt_world_load_obj_ex(w3._h, 0, str(path).encode(), 1.0, (C.c_float * 3)(0, 0, 0), 0, 4) == b200rt.ERR_INVALID
    bad = tmp_path / "badvt.obj"
    bad.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1/1 2/1 3/1\n")            # vt index without any vt
    with pytest.raises(b200rt.B200rtError):
        o3.load_obj_ex(str(bad))


def test_obj_errors(b200rt, tmp_path):
    w = b200rt.World()
    o = w.push_object(b200rt.color_material())
    with pytest.raises(b200rt.B200rtError) as e:               # the reference asserts / panics here (main.rs:785)
        o.load_obj(str(tmp_path / "missing.obj"))
    assert e.value.code == b200rt.ERR_IO
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nf 1 2 3\n")                        # index out of range
    with pytest.raises(b200rt.B200rtError):
        o.load_obj(str(bad))
    empty = tmp_path / "empty.obj"
    empty.write_text("# nothing\n")
    with pytest.raises(b200rt.B200rtError):
        o.load_obj(str(empty))
    lib = b200rt.load_library()
    assert lib.b200rt_world_push_sphere(w._h, 99, (C.c_float * 3)(0, 0, 0), 1.0) == b200rt.ERR_INVALID   # unknown object


def test_golden_fixture_scene_equals_the_builder(b200rt, fixture_world):
    """tests/golden/fixture_scene.npz (what bench.py's reference arm renders, without loading libb200rt.so) is the
    scene literal of main.rs:810-1075 as World.fixture() builds it, byte for byte; same camera and defaults."""
    import ctypes as C
    import oracle_binding as ob
    g = ob.GoldenFixture()
    s, t = fixture_world.scene(), g.scene
    for n_field, ptr_field, typ in (("n_triangles", "triangles", b200rt.Triangle), ("n_spheres", "spheres", b200rt.Sphere),
                                    ("n_materials", "materials", b200rt.Material), ("n_lights", "lights", b200rt.Light)):
        n = getattr(s, n_field)
        assert n == getattr(t, n_field), n_field
        assert C.string_at(getattr(s, ptr_field), n * C.sizeof(typ)) == C.string_at(getattr(t, ptr_field), n * C.sizeof(typ)), ptr_field
    assert bytes(g.camera) == bytes(b200rt.fixture_camera())
    assert bytes(g.params()) == bytes(b200rt.default_params())
    assert g.params(width=64).width == 64
