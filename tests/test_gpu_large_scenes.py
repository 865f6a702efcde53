"""Scenes that span many 64-triangle tiles (C5: fixture + synthetic OBJ mesh) and the 4K configurations (C3/C4),
checked through size-independent properties where the oracle would take too long at full size."""
import numpy as np
import pytest

import scene_util
from test_gpu_parity import assert_hits_equal, assert_stochastic_agreement, check_whitted, random_rays, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mesh_ctx(b200rt, tmp_path_factory):
    world, ntri = scene_util.fixture_plus_mesh(b200rt, tmp_path_factory.mktemp("mesh"), n=41)   # 3200 + 64 triangles
    ctx = b200rt.Context(0)
    ctx.upload_scene(world)
    yield ctx, world, ntri
    ctx.close()


def test_multi_tile_intersect_bit_exact(b200rt, oracle, mesh_ctx):
    ctx, world, ntri = mesh_ctx
    assert world.scene().n_triangles == 64 + ntri and ntri == 3200       # 51 tiles, the last one ragged
    rays = random_rays(b200rt, 1 << 18, 99)
    rays["exclude_prim"] = np.where(rays["exclude_prim"] >= 0, rays["exclude_prim"] * 47 % (64 + ntri + 4), -1)
    g = ctx.intersect(rays, b200rt.CAST_TWO_PHASE)
    o = oracle.intersect(world.scene(), rays)
    assert (o["prim_id"] >= 64).mean() > 0.02                            # the mesh is hit
    assert_hits_equal(g, o)
    gb = ctx.intersect(rays[: 1 << 14], b200rt.CAST_BRUTE_EXACT)
    assert_hits_equal(gb, o[: 1 << 14])


def test_c5_shape_whitted_and_samples(b200rt, oracle, mesh_ctx):
    """C5 shape at reduced size: Whitted ids + colours, and one stochastic sample per pixel ("photons",
    PhotonAccumulator semantics) against the oracle."""
    ctx, world, _ = mesh_ctx
    params = b200rt.default_params(width=200, height=125)
    check_whitted(b200rt, oracle, ctx, world, params)
    cam = b200rt.fixture_camera()
    acc = ctx.render_distributed(cam, params, 0, 1)
    o_acc, _ = oracle.render_distributed(world.scene(), cam, params, 0, 1)
    assert_stochastic_agreement(acc, o_acc, "C5 shape 200x125 x 1 epoch, 3264 triangles", 1, min_psnr=35.0)


def test_c5_full_size_pixel_sample(b200rt, oracle, tmp_path_factory):
    """C5 at its REAL size: fixture scene + the 100 352-triangle OBJ height field (1570 tiles), 4000x2500,
    one stochastic sample per pixel.  The oracle renders 2000 pixels of the frame (40 rows x 50 columns, one sample
    each: ~0.7 M ray x primitive tests per pixel sample on the CPU); the GPU renders those rows of the full frame through
    the tiled TMA cast and must produce the same samples: same accepted set, colours to libm ulps."""
    from concurrent.futures import ThreadPoolExecutor
    world, ntri = scene_util.fixture_plus_mesh(b200rt, tmp_path_factory.mktemp("mesh_full"), n=225)
    assert ntri == 100352
    ctx = b200rt.Context(0)
    ctx.upload_scene(world)
    cam = b200rt.fixture_camera()
    W, H = 4000, 2500
    params = b200rt.default_params(width=W, height=H, seed=0)
    rng = np.random.default_rng(2025)
    rows = np.sort(rng.choice(np.arange(700, 2300), size=40, replace=False))      # rows that see the mesh and the room
    cols = np.sort(rng.choice(W, size=50, replace=False))
    scene = world.scene()

    def one(yx):
        return oracle.sample_distributed(scene, cam, params, int(yx[0]), int(yx[1]), 0)

    todo = [(y, x) for y in rows for x in cols]
    with ThreadPoolExecutor(max_workers=oracle.host_threads()) as pool:          # ctypes drops the GIL
        o = np.array(list(pool.map(one, todo)), dtype=np.float32).reshape(len(rows), len(cols), 3)
    g = np.zeros((len(rows), len(cols), 4), dtype=np.float32)
    for i, y in enumerate(rows):
        acc = ctx.render_distributed(cam, b200rt.copy_params(params, row_begin=int(y), row_count=1), 0, 1)
        g[i] = acc[y, cols]
    normal = np.isfinite(o).all(axis=2) & ((np.abs(o) >= np.finfo(np.float32).tiny) & (np.abs(o) < np.inf)).all(axis=2)
    assert np.array_equal(g[..., 3] == 1.0, normal)                       # the is_normal filter accepts the same samples
    assert normal.mean() > 0.5
    err = np.abs(g[..., :3] - o) / np.maximum(np.maximum(np.abs(o), np.abs(g[..., :3])), 1e-3)
    err = np.where(normal[..., None], err, 0.0)
    moved = (err.max(axis=2) > 1e-3).sum()
    print(f"C5 full size: {normal.sum()} accepted of {normal.size} samples, {moved} moved, max rel err of the others "
          f"{np.where(err.max(axis=2) > 1e-3, 0.0, err.max(axis=2)).max():.2e}")
    assert moved <= 4, moved
    ctx.close()


def test_ragged_triangle_counts(b200rt, oracle):
    """Tile edge cases: 1, 31, 32, 33, 63, 65, 128, 129 triangles (+ spheres only, + empty scene)."""
    rng = np.random.default_rng(5)
    for nt in (0, 1, 31, 32, 33, 63, 65, 128, 129):
        w = b200rt.World()
        o = w.push_object(b200rt.color_material(diffuse_color=(0.8, 0.7, 0.6), shiness=0.3))
        for _ in range(nt):
            c = rng.uniform(-1.5, 1.5, size=3)
            o.push_flat_triangle((c + rng.uniform(-0.4, 0.4, size=(3, 3))).astype(np.float32))
        if nt % 2 == 1 or nt == 0:
            o.push_sphere([0.2, 0.1, -0.3], 0.5)
        w.push_light(b200rt.point_light([0, 3, 0], [1, 1, 1]))
        ctx = b200rt.Context(0)
        ctx.upload_scene(w)
        rays = random_rays(b200rt, 1 << 14, 7 + nt)
        rays["exclude_prim"] = np.where(rays["exclude_prim"] >= 0, rays["exclude_prim"] % max(nt + 1, 1), -1)
        assert_hits_equal(ctx.intersect(rays), oracle.intersect(w.scene(), rays))
        ctx.close()


def test_multi_tile_rays_outside_the_filter_assumptions(b200rt, oracle, mesh_ctx):
    """NaN components (the walk ends with the last admissible triangle: cast_nan_ray_triangles), origins at infinity /
    beyond the packed bound and non-unit directions (the warp tests the tile together and folds the nearest-so-far rule
    in index order: rl_coop_exact_tile), mixed with ordinary rays in the same warps, on a 51-tile scene: ids, faces and
    distances as the reference's ordered walk over every triangle, through both cast kernels."""
    ctx, world, ntri = mesh_ctx
    n = 4096
    rays = random_rays(b200rt, n, 7)
    rng = np.random.default_rng(8)
    kind = rng.integers(0, 8, size=n)                        # 0..3 ordinary
    o, d = rays["origin"].copy(), rays["direction"].copy()
    sel = kind == 4; o[sel, rng.integers(0, 3, size=sel.sum())] = np.nan
    sel = kind == 5; d[sel, rng.integers(0, 3, size=sel.sum())] = np.nan
    sel = kind == 6                                           # origin at +-inf on one axis, or merely far away
    ax = rng.integers(0, 3, size=sel.sum())
    far = np.where(rng.random(sel.sum()) < 0.5, 3e4, -3e4)
    o[sel, ax] = np.where(rng.random(sel.sum()) < 0.7, np.where(far > 0, np.inf, -np.inf), far)
    sel = kind == 7; d[sel] *= rng.uniform(0.2, 5.0, size=(sel.sum(), 1)).astype(np.float32)   # |dir| != 1
    rays["origin"], rays["direction"] = o, d
    ob = oracle.intersect(world.scene(), rays)
    for mode in ("rl", "transposed"):
        import os
        if mode == "transposed":
            os.environ["B200RT_INTERSECT"] = "transposed"
        try:
            g = ctx.intersect(rays, b200rt.CAST_TWO_PHASE)
        finally:
            os.environ.pop("B200RT_INTERSECT", None)
        assert np.array_equal(g["prim_id"], ob["prim_id"]), (mode, int((g["prim_id"] != ob["prim_id"]).sum()))
        hit = ob["prim_id"] >= 0
        assert np.array_equal(g["face_direction"][hit], ob["face_direction"][hit]), mode
        gd, od = g["distance"][hit], ob["distance"][hit]
        assert np.array_equal(np.isnan(gd), np.isnan(od)), mode
        fin = ~np.isnan(od)
        assert np.array_equal(gd[fin].view(np.uint32), od[fin].view(np.uint32)), mode
    assert (kind >= 4).sum() > 1500 and (ob["prim_id"][kind >= 4] >= 0).sum() > 100    # the odd rays do "hit" things


def test_degenerate_and_far_inputs_bypass_the_filter(b200rt, oracle):
    """Zero-area triangles (NaN normal in the reference), rays far outside the packed origin bound, non-unit and
    non-finite directions: the conservative filter must hand all of them to the exact test."""
    w = b200rt.World()
    o = w.push_object(b200rt.color_material())
    o.push_flat_triangle([[0, 0, 0], [1, 0, 0], [0, 1, 0]])
    o.push_flat_triangle([[0, 0, 1], [0, 0, 1], [0, 0, 1]])            # degenerate: face_normal = NaN
    o.push_flat_triangle([[0, 0, -1], [1, 0, -1], [0, 1, -1]])
    ctx = b200rt.Context(0)
    ctx.upload_scene(w)
    rays = np.zeros(6, dtype=b200rt.RAY_DTYPE)
    rays["exclude_prim"] = -1
    rays["origin"] = [[0.2, 0.2, 5], [0.2, 0.2, 5e4], [0.2, 0.2, 5], [0.2, 0.2, -5], [0.2, 0.2, 5], [np.nan, 0, 0]]
    rays["direction"] = [[0, 0, -1], [0, 0, -1], [0, 0, -3], [0, 0, 1], [np.inf, 0, -1], [0, 0, -1]]
    rays["face_direction"] = [0, 0, 0, 1, 0, 2]
    g = ctx.intersect(rays)
    ob = oracle.intersect(w.scene(), rays)
    assert np.array_equal(g["prim_id"], ob["prim_id"]), (g["prim_id"], ob["prim_id"])
    hit = ob["prim_id"] >= 0
    gd, od = g["distance"][hit], ob["distance"][hit]
    assert np.array_equal(np.isnan(gd), np.isnan(od))                  # NaN "hits" of the reference are reproduced
    fin = ~np.isnan(od)                                                # (NaN payload bits differ between x86 and sm_100)
    assert np.array_equal(gd[fin].view(np.uint32), od[fin].view(np.uint32))
    ctx.close()


def test_c3_4k_properties(b200rt, oracle, gpu_ctx, fixture_world):
    """C3 (3840x2160 Whitted, 1 GPU): deterministic, row bands bitwise equal to the full frame, and a band of rows
    equal to the oracle (ids + 1e-4)."""
    cam = b200rt.fixture_camera()
    params = b200rt.default_params(width=3840, height=2160)
    full, prim = gpu_ctx.render_whitted(cam, params)
    again, prim2 = gpu_ctx.render_whitted(cam, params)
    assert np.array_equal(full.view(np.uint32), again.view(np.uint32)) and np.array_equal(prim, prim2)
    bad = ~np.isfinite(full).all(axis=2)
    assert bad.sum() <= 8                                               # the reference's own NaN pixels (1 at this size)
    for y in sorted(set(np.where(bad)[0].tolist()))[:2]:                # ... are NaN in the oracle too, same positions
        check_whitted(b200rt, oracle, gpu_ctx, fixture_world, b200rt.copy_params(params, row_begin=int(y), row_count=1))
    assert (prim >= 0).mean() > 0.5, (prim >= 0).mean()                 # 16:9 frame: more background than 4:3
    band = b200rt.copy_params(params, row_begin=1000, row_count=96)
    check_whitted(b200rt, oracle, gpu_ctx, fixture_world, band)
    out = np.zeros_like(full)
    gpu_ctx.render_whitted(cam, band, out_rgb=out, want_prim_id=False)
    assert np.array_equal(out[1000:1096].view(np.uint32), full[1000:1096].view(np.uint32))
    assert not out[:1000].any() and not out[1096:].any()               # only the requested rows are written


def test_c4_4k_epoch_properties(b200rt, oracle, gpu_ctx, fixture_world):
    """C4 shape (3840x2160 stochastic): epoch ranges add (2 GPUs' worth of shards == one launch), counts are
    bounded by the epoch count, and a row band matches the oracle sample for sample."""
    cam = b200rt.fixture_camera()
    params = b200rt.default_params(width=3840, height=2160, seed=0)
    a = gpu_ctx.render_distributed(cam, params, 0, 4)
    bsum = gpu_ctx.render_distributed(cam, params, 0, 2)
    gpu_ctx.render_distributed(cam, params, 2, 2, bsum)
    assert np.array_equal(a[..., 3], bsum[..., 3]) and a[..., 3].max() == 4
    np.testing.assert_allclose(bsum[..., :3], a[..., :3], rtol=2e-6, atol=1e-7)
    band = b200rt.copy_params(params, row_begin=1040, row_count=48)
    g = gpu_ctx.render_distributed(cam, band, 0, 4)
    o_acc, _ = oracle.render_distributed(fixture_world.scene(), cam, band, 0, 4)
    assert np.array_equal(g[1040:1088, :, 3], o_acc[1040:1088, :, 3])
    assert np.array_equal(g[1040:1088].view(np.uint32), a[1040:1088].view(np.uint32))     # band == same rows of the full frame
    assert_stochastic_agreement(g[1040:1088], o_acc[1040:1088], "C4 band 3840x48 x 4 epochs", 4)


def test_split_tile_range_cast_bitwise(b200rt, mesh_ctx):
    """Rounds of few rays on a scene of many tiles take the split form of the tiled cast (a 128-ray block per CTA, the tile
    range over its four warps, partial results folded in tile order): the same bits as the unsplit cast, for trusted rays,
    rays from infinity and NaN rays alike (the stochastic tracer produces all three on this scene)."""
    import os
    ctx, world, _ = mesh_ctx
    cam = b200rt.fixture_camera()
    p = b200rt.default_params(width=400, height=250, seed=9)
    out = {}
    for name, sb in (("split", None), ("unsplit", "0"), ("always", "1000000")):
        if sb is None:
            os.environ.pop("B200RT_WF_SPLIT_BELOW", None)
        else:
            os.environ["B200RT_WF_SPLIT_BELOW"] = sb
        try:
            ctx.reset_stats()
            out[name] = (ctx.render_distributed(cam, p, 0, 2), ctx.stats())
        finally:
            os.environ.pop("B200RT_WF_SPLIT_BELOW", None)
    ref = out["unsplit"][0]
    assert (ref[..., 3] > 0).mean() > 0.5
    for name in ("split", "always"):
        assert np.array_equal(out[name][0].view(np.uint32), ref.view(np.uint32)), name
        assert out[name][1]["casts"] == out["unsplit"][1]["casts"]


def test_device_round_loop_matches_host_rounds(b200rt, gpu_ctx):
    """The wavefront's rounds repeated on the device (CUDA graph WHILE node) and enqueued from the host: the same bits, the
    same casts; the device loop reports its rounds and launches through the device counters."""
    import os
    cam = b200rt.fixture_camera()
    p = b200rt.default_params(width=640, height=360, seed=21)
    gpu_ctx.reset_stats()
    a = gpu_ctx.render_distributed(cam, p, 0, 5)
    sa = gpu_ctx.stats()
    os.environ["B200RT_WF_GRAPH"] = "0"
    try:
        gpu_ctx.reset_stats()
        h = gpu_ctx.render_distributed(cam, p, 0, 5)
        sh = gpu_ctx.stats()
    finally:
        del os.environ["B200RT_WF_GRAPH"]
    assert np.array_equal(a.view(np.uint32), h.view(np.uint32))
    assert sa["casts"] == sh["casts"] and sa["samples"] == sh["samples"]
    assert sa["wavefront_rounds"] >= 15 and sa["wavefront_rounds"] % 2 == 1        # round 0 + pairs of rounds
    assert sa["kernel_launches"] > 5 * sa["wavefront_rounds"] // 2
