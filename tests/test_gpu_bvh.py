"""World::cast through the acceleration structure (B200RT_CAST_BVH, csrc/rt_bvh.cuh) returns the hits of the
reference's walk over every primitive, bit for bit - against the oracle where it finishes in seconds, and against the
two-phase / brute-force casts (which tests/test_gpu_parity.py and test_gpu_large_scenes.py pin to the oracle) elsewhere."""
import numpy as np
import pytest

import scene_util
from test_gpu_parity import assert_hits_equal, random_rays

pytestmark = pytest.mark.gpu


def assert_same_bits(g, r):
    """all fields of two GPU casts (same libm on both sides: normals and uv are bitwise too)"""
    assert np.array_equal(g["prim_id"], r["prim_id"]), f"{int((g['prim_id'] != r['prim_id']).sum())} ids differ"
    for f in ("face_direction", "object_index"):
        assert np.array_equal(g[f], r[f]), f
    for f in ("distance", "position", "normal", "uv"):
        assert np.array_equal(g[f].view(np.uint32), r[f].view(np.uint32)), f


def test_bvh_fixture_random_rays_bit_exact(b200rt, oracle, gpu_ctx, fixture_world):
    rays = random_rays(b200rt, 1 << 20, 4321)
    g = gpu_ctx.intersect(rays, b200rt.CAST_BVH)
    o = oracle.intersect(fixture_world.scene(), rays)
    assert_hits_equal(g, o)
    assert_same_bits(g, gpu_ctx.intersect(rays, b200rt.CAST_TWO_PHASE))


def test_bvh_grazing_rays_hit_at_infinity(b200rt, oracle, gpu_ctx, fixture_world):
    """the rays of test_grazing_rays_hit_at_infinity (n.dir == 0 exactly for one triangle: t = +inf hits, main.rs:204-231)
    and rays lying IN the plane of a triangle (0/0: NaN distances, the order-dependent walk)"""
    cases = [("3f959753 00000000 bf87819a", "bf00d52e 3f3c2e41 bee89aa4", 0, 37, 1),
             ("40042e8a 4024e36a 400203d6", "befd3bb2 bead9492 bf4cdeb7", 0, -1, 0),
             ("3f83e17c 00000000 bf0a1ae0", "beaf1a07 3f4ca2b9 befcf18c", 0, 37, 1),
             ("bfd6e906 00000000 3fb82ed9", "3f7b6335 3e249b5d bdcb772a", 0, 36, 1)]
    rays = np.zeros(64 + 4096, dtype=b200rt.RAY_DTYPE)
    rays["exclude_prim"] = -1
    for i in range(64):
        o, d, face, ex, exf = cases[i % len(cases)]
        rays["origin"][i] = np.array([int(x, 16) for x in o.split()], dtype=np.uint32).view(np.float32)
        rays["direction"][i] = np.array([int(x, 16) for x in d.split()], dtype=np.uint32).view(np.float32)
        rays["face_direction"][i] = face; rays["exclude_prim"][i] = ex; rays["exclude_face"][i] = exf
    # rays inside the floor plane y = 0 and along the axis-aligned walls / slab faces (exact zeros in n.dir and d - n.o)
    rng = np.random.default_rng(3)
    k = np.arange(64, len(rays))
    axis = rng.integers(0, 3, size=len(k))
    o = rng.uniform(-1.9, 1.9, size=(len(k), 3)).astype(np.float32)
    plane = rng.choice(np.array([0.0, 2.0, -2.0, 1.0, 1.5, 0.6, 0.7, 0.71, 0.81, 0.5, -0.5, 0.3, -0.3], dtype=np.float32), size=len(k))
    o[np.arange(len(k)), axis] = plane
    d = rng.normal(size=(len(k), 3)).astype(np.float32)
    d[np.arange(len(k)), axis] = 0.0
    # half of them along a coordinate axis inside the plane
    ax2 = (axis + 1 + rng.integers(0, 2, size=len(k))) % 3
    snap = rng.random(len(k)) < 0.5
    d[snap] = 0.0
    d[snap, ax2[snap]] = np.where(rng.random(snap.sum()) < 0.5, 1.0, -1.0)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["origin"][k] = o
    rays["direction"][k] = d.astype(np.float32)
    rays["face_direction"][k] = rng.integers(0, 3, size=len(k))
    ref = oracle.intersect(fixture_world.scene(), rays)
    g = gpu_ctx.intersect(rays, b200rt.CAST_BVH)
    assert np.array_equal(g["prim_id"], ref["prim_id"]), np.where(g["prim_id"] != ref["prim_id"])[0][:10]
    hit = ref["prim_id"] >= 0
    gd, od = g["distance"][hit], ref["distance"][hit]
    assert np.array_equal(np.isnan(gd), np.isnan(od))
    fin = ~np.isnan(od)
    assert np.array_equal(gd[fin].view(np.uint32), od[fin].view(np.uint32))
    assert np.isinf(od).sum() >= 16 and np.all(np.isinf(g["distance"][:64]))
    assert_same_bits(g, gpu_ctx.intersect(rays, b200rt.CAST_TWO_PHASE))


@pytest.fixture(scope="module")
def mesh_ctx(b200rt, tmp_path_factory):
    world, ntri = scene_util.fixture_plus_mesh(b200rt, tmp_path_factory.mktemp("mesh_bvh"), n=97)   # 18 432 + 64 triangles
    ctx = b200rt.Context(0)
    ctx.upload_scene(world)
    yield ctx, world, ntri
    ctx.close()


def adversarial_rays(b200rt, n, seed, n_prims):
    """random rays + rays outside the filter's assumptions: NaN / inf components, far origins, non-unit directions,
    axis-parallel directions, origins on the mesh"""
    rays = random_rays(b200rt, n, seed)
    rng = np.random.default_rng(seed + 1)
    rays["exclude_prim"] = np.where(rays["exclude_prim"] >= 0, rng.integers(0, n_prims, size=n), -1)
    idx = rng.permutation(n)
    q = n // 64
    rays["origin"][idx[0:q], 0] = np.nan
    rays["direction"][idx[q:2 * q], 1] = np.nan
    rays["origin"][idx[2 * q:3 * q]] *= 1e5                      # beyond the packed origin bound
    rays["origin"][idx[3 * q:4 * q], 2] = np.inf                 # rays from infinity (after a t = +inf hit)
    rays["direction"][idx[4 * q:5 * q]] *= 3.0                   # |dir| != 1
    d = np.zeros((q, 3), dtype=np.float32)
    d[np.arange(q), rng.integers(0, 3, size=q)] = rng.choice(np.array([-1.0, 1.0], dtype=np.float32), size=q)
    rays["direction"][idx[5 * q:6 * q]] = d                      # axis-parallel: exact zeros in n.dir for the room's faces
    rays["direction"][idx[6 * q:7 * q], 2] = 0.0                 # (not renormalised: slightly short directions)
    return rays


def test_bvh_mesh_rays_bit_exact(b200rt, oracle, mesh_ctx):
    ctx, world, ntri = mesh_ctx
    n_prims = 64 + ntri + 4
    rays = adversarial_rays(b200rt, 1 << 19, 77, n_prims)
    g = ctx.intersect(rays, b200rt.CAST_BVH)
    st = ctx.stats()
    t = ctx.intersect(rays, b200rt.CAST_TWO_PHASE)
    assert (t["prim_id"] >= 64).mean() > 0.02 and (t["prim_id"] >= 64 + ntri).sum() > 0      # mesh and spheres are hit
    nan = np.isnan(t["distance"])
    assert np.array_equal(np.isnan(g["distance"]), nan)
    for f in ("prim_id", "face_direction", "object_index"):
        assert np.array_equal(g[f], t[f]), f
    for f in ("distance", "position", "normal", "uv"):
        a, b = g[f].view(np.uint32).reshape(len(g), -1), t[f].view(np.uint32).reshape(len(g), -1)
        assert np.array_equal(a[~nan], b[~nan]), f
    # ... and a sample of them against the oracle itself
    sub = rays[: 1 << 13]
    o = oracle.intersect(world.scene(), sub)
    assert np.array_equal(g["prim_id"][: 1 << 13], o["prim_id"])
    ok = (o["prim_id"] >= 0) & ~np.isnan(o["distance"])
    assert np.array_equal(g["distance"][: 1 << 13][ok].view(np.uint32), o["distance"][ok].view(np.uint32))
    # the structure is doing its job: far fewer exact tests than ray x triangle pairs
    trusted = len(rays) - 7 * (len(rays) // 64)
    print(f"BVH: {st['exact_confirms'] / len(rays):.1f} exact tests per cast ({64 + ntri} triangles), {st['certify_fallbacks']} ordered walks")
    assert st["exact_confirms"] < 0.2 * trusted * (64 + ntri)


def test_bvh_frames_bitwise(b200rt, mesh_ctx):
    """Whitted and stochastic frames of the mesh scene: the acceleration structure behind both tracers and both schedules
    of the stochastic one gives the two-phase cast's bits."""
    ctx, world, _ = mesh_ctx
    cam = b200rt.fixture_camera()
    pw = b200rt.default_params(width=320, height=200)
    rgb, prim = ctx.render_whitted(cam, pw)
    rgb_b, prim_b = ctx.render_whitted(cam, b200rt.copy_params(pw, cast_mode=b200rt.CAST_BVH))
    assert np.array_equal(prim, prim_b) and np.array_equal(rgb.view(np.uint32), rgb_b.view(np.uint32))
    assert (prim >= 64).mean() > 0.05
    ref = ctx.render_distributed(cam, b200rt.default_params(width=320, height=200, seed=5), 0, 3)
    for tracer in (b200rt.TRACER_WAVEFRONT, b200rt.TRACER_MEGAKERNEL):
        p = b200rt.default_params(width=320, height=200, seed=5, tracer=tracer, cast_mode=b200rt.CAST_BVH)
        acc = ctx.render_distributed(cam, p, 0, 3)
        assert np.array_equal(acc[..., 3], ref[..., 3]), tracer
        if tracer == b200rt.TRACER_WAVEFRONT:
            assert np.array_equal(acc.view(np.uint32), ref.view(np.uint32))
        else:   # (the megakernel sums a pixel's epochs in one thread, the wavefront per slot: fp32 summation order)
            np.testing.assert_allclose(acc[..., :3], ref[..., :3], rtol=2e-6, atol=1e-7)


def test_bvh_ragged_and_degenerate_scenes(b200rt, oracle):
    rng = np.random.default_rng(11)
    for nt in (0, 1, 5, 33, 129):
        w = b200rt.World()
        o = w.push_object(b200rt.color_material(diffuse_color=(0.8, 0.7, 0.6), shiness=0.3))
        for _ in range(nt):
            c = rng.uniform(-1.5, 1.5, size=3)
            o.push_flat_triangle((c + rng.uniform(-0.4, 0.4, size=(3, 3))).astype(np.float32))
        if nt >= 5:
            o.push_flat_triangle([[0, 0, 1], [0, 0, 1], [0, 0, 1]])          # zero area: NaN normal in the reference
            o.push_flat_triangle([[0, 0, 0], [1, 0, 0], [1, 1e-7, 0]])       # a needle
        if nt % 2 == 1 or nt == 0:
            o.push_sphere([0.2, 0.1, -0.3], 0.5)
        w.push_light(b200rt.point_light([0, 3, 0], [1, 1, 1]))
        ctx = b200rt.Context(0)
        ctx.upload_scene(w)
        rays = random_rays(b200rt, 1 << 14, 70 + nt)
        rays["exclude_prim"] = np.where(rays["exclude_prim"] >= 0, rays["exclude_prim"] % (nt + 3), -1)
        g = ctx.intersect(rays, b200rt.CAST_BVH)
        ob = oracle.intersect(w.scene(), rays)
        assert np.array_equal(g["prim_id"], ob["prim_id"]), nt
        hit = (ob["prim_id"] >= 0) & ~np.isnan(ob["distance"])
        assert np.array_equal(g["distance"][hit].view(np.uint32), ob["distance"][hit].view(np.uint32)), nt
        ctx.close()
