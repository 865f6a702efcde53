"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star): deterministic scenes — identical hit-primitive ids and
|a-b| <= 1e-4 * max(|a|,|b|,1e-3) per channel; stochastic scenes — converged-mean PSNR >= 40 dB at
equal sample count (and, because oracle and GPU share the Philox sample stream, near-identical
individual samples)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4
ABS_FLOOR = 1e-3


def rel_err(a, b):
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), ABS_FLOOR)


def check_whitted(b200rt, oracle, ctx, world, params, cam=None):
    cam = cam or b200rt.fixture_camera()
    rgb, prim = ctx.render_whitted(cam, params)
    o_rgb, o_prim, cnt = oracle.render_whitted(world.scene(), cam, params)
    r0 = params.row_begin if params.row_count else 0
    r1 = r0 + (params.row_count if params.row_count else params.height)
    assert np.array_equal(prim[r0:r1], o_prim[r0:r1]), f"{int((prim[r0:r1] != o_prim[r0:r1]).sum())} hit ids differ"
    # the reference itself yields NaN on isolated pixels (e.g. one on the glass sphere at 3840x2160): they must
    # be reproduced at the same positions; everything else is compared numerically
    fin_g, fin_o = np.isfinite(rgb[r0:r1]), np.isfinite(o_rgb[r0:r1])
    assert np.array_equal(fin_g, fin_o), f"{int((fin_g != fin_o).sum())} non-finite channel positions differ"
    assert (~fin_o).sum() <= max(24, 1e-4 * fin_o.size)
    err = np.where(fin_o, rel_err(np.where(fin_g, rgb[r0:r1], 0), np.where(fin_o, o_rgb[r0:r1], 0)), 0.0)
    assert err.max() <= REL_TOL, f"max rel err {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"
    return rgb, prim, cnt


@pytest.mark.parametrize("mode", ["two_phase", "brute_exact"])
def test_c1_whitted_fixture_1280x960(b200rt, oracle, gpu_ctx, fixture_world, mode):
    """C1: the reference's own frame (main.rs:1084-1101): 1280x960, depth 5."""
    params = b200rt.default_params(cast_mode=b200rt.CAST_TWO_PHASE if mode == "two_phase" else b200rt.CAST_BRUTE_EXACT)
    gpu_ctx.reset_stats()
    rgb, prim, cnt = check_whitted(b200rt, oracle, gpu_ctx, fixture_world, params)
    st = gpu_ctx.stats()
    assert st["casts"] == cnt["casts"], (st, cnt)        # same number of World::cast calls as the reference recursion
    assert st["samples"] == cnt["samples"] == 1280 * 960
    assert (prim >= 0).sum() > 1_000_000                 # the frame is mostly covered (144k black px in the reference image)


def test_c2_whitted_depth8_1920x1080(b200rt, oracle, gpu_ctx, fixture_world):
    params = b200rt.default_params(width=1920, height=1080, depth=8)
    check_whitted(b200rt, oracle, gpu_ctx, fixture_world, params)


@pytest.mark.parametrize("depth", [0, 1, 2])
def test_whitted_shallow_depths(b200rt, oracle, gpu_ctx, fixture_world, depth):
    params = b200rt.default_params(width=320, height=240, depth=depth)
    check_whitted(b200rt, oracle, gpu_ctx, fixture_world, params)


def test_whitted_ragged_sizes(b200rt, oracle, gpu_ctx, fixture_world):
    for (w, h) in [(1, 1), (17, 9), (33, 65), (250, 3)]:
        params = b200rt.default_params(width=w, height=h)
        check_whitted(b200rt, oracle, gpu_ctx, fixture_world, params)


def random_rays(b200rt, n, seed, world_scene=None):
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, dtype=b200rt.RAY_DTYPE)
    o = rng.uniform(-2.5, 2.5, size=(n, 3)).astype(np.float32)
    o[:, 1] = rng.uniform(-0.5, 3.0, size=n).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    rays["origin"] = o
    rays["direction"] = d.astype(np.float32)
    rays["face_direction"] = rng.integers(0, 3, size=n)
    ex = rng.integers(-1, 68, size=n)
    rays["exclude_prim"] = np.where(rng.random(n) < 0.5, -1, ex)
    rays["exclude_face"] = rng.integers(0, 3, size=n)
    return rays


def assert_hits_equal(g, o):
    assert np.array_equal(g["prim_id"], o["prim_id"]), f"{int((g['prim_id'] != o['prim_id']).sum())} ids differ"
    hit = o["prim_id"] >= 0
    assert np.array_equal(g["face_direction"][hit], o["face_direction"][hit])
    assert np.array_equal(g["object_index"][hit], o["object_index"][hit])
    # distance and position are pure +,-,*,/,sqrt arithmetic: bit-identical
    assert np.array_equal(g["distance"][hit].view(np.uint32), o["distance"][hit].view(np.uint32))
    assert np.array_equal(g["position"][hit].view(np.uint32), o["position"][hit].view(np.uint32))
    np.testing.assert_allclose(g["normal"][hit], o["normal"][hit], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(g["uv"][hit], o["uv"][hit], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("mode", ["two_phase", "brute_exact"])
def test_intersect_random_rays_bit_exact(b200rt, oracle, gpu_ctx, fixture_world, mode):
    """World::cast on 1M random rays (all face modes, random exclusions)."""
    rays = random_rays(b200rt, 1 << 20, 1234)
    g = gpu_ctx.intersect(rays, b200rt.CAST_TWO_PHASE if mode == "two_phase" else b200rt.CAST_BRUTE_EXACT)
    o = oracle.intersect(fixture_world.scene(), rays)
    assert (o["prim_id"] >= 0).mean() > 0.2
    assert_hits_equal(g, o)


def test_intersect_empty(b200rt, gpu_ctx):
    rays = np.zeros(0, dtype=b200rt.RAY_DTYPE)
    assert gpu_ctx.intersect(rays).shape == (0,)


# Measured on B200 (round 2, 1280x960 x 8 epochs = 9.8 M samples): 22 pixels with a moved sample (1.8e-5 of the pixels),
# 5 pixels whose accepted-sample count differs (a libm ulp makes a sample exactly 0, which the is_normal filter of
# main.rs:1157-1160 drops), PSNR of the mean images 82 dB.  The bounds leave a factor of ~10 for other CUDA / glibc libm
# versions; a path that really differs shows up as 100x these.
MOVED_BOUND = 2e-4
COUNT_DIFF_BOUND = 4e-5


def assert_stochastic_agreement(acc, o_acc, what, epochs, min_psnr=40.0):
    from parity_util import stochastic_agreement
    s = stochastic_agreement(acc, o_acc)
    print(f"{what}: {s}")
    n_px = int(np.prod(np.asarray(acc).shape[:-1]))
    assert s["count_diff"] <= max(2, int(COUNT_DIFF_BOUND * n_px * epochs)), s   # the is_normal filter accepts the same samples
    assert s["nonfinite_mismatch"] == 0, s
    assert s["frac_moved"] <= max(MOVED_BOUND, 2.0 / n_px), s   # pixels in which a libm ulp moved a sample across a silhouette
    assert s["max_abs"] <= 8.0 * s["peak"] * epochs + 1e-6 or s["frac_moved"] == 0.0, s   # ... and by no more than samples can
    assert s["psnr"] >= min_psnr, s                     # north_star: converged-mean PSNR >= 40 dB at equal sample count
    return s


def test_distributed_samples_match_oracle(b200rt, oracle, gpu_ctx, fixture_world):
    """C4 semantics at small size: 4 epochs, per-pixel {sum.rgb, count} against the oracle."""
    cam = b200rt.fixture_camera()
    params = b200rt.default_params(width=320, height=240, seed=7)
    acc = gpu_ctx.render_distributed(cam, params, 0, 4)
    o_acc, cnt = oracle.render_distributed(fixture_world.scene(), cam, params, 0, 4)
    assert_stochastic_agreement(acc, o_acc, "320x240 x 4 epochs", 4)


def test_distributed_reference_frame_psnr(b200rt, oracle, gpu_ctx, fixture_world):
    """The reference's own frame size (1280x960, main.rs:1084-1085), 8 epochs: PSNR of the mean images and the
    fraction of moved samples at a size where a pixel sees more silhouettes than at 320x240."""
    cam = b200rt.fixture_camera()
    params = b200rt.default_params(width=1280, height=960, seed=11)
    acc = gpu_ctx.render_distributed(cam, params, 0, 8)
    o_acc, cnt = oracle.render_distributed(fixture_world.scene(), cam, params, 0, 8, n_threads=oracle.host_threads())
    assert_stochastic_agreement(acc, o_acc, "1280x960 x 8 epochs", 8)


def test_distributed_epoch_split_additive(b200rt, gpu_ctx):
    """Epoch sharding: rendering [0,6) equals rendering [0,2)+[2,4)+[4,6) into the same buffer (fp32 sum order)."""
    cam = b200rt.fixture_camera()
    params = b200rt.default_params(width=256, height=192, seed=3)
    full = gpu_ctx.render_distributed(cam, params, 0, 6)
    parts = np.zeros_like(full)
    for e in (0, 2, 4):
        gpu_ctx.render_distributed(cam, params, e, 2, parts)
    assert np.array_equal(full[..., 3], parts[..., 3])
    np.testing.assert_allclose(parts[..., :3], full[..., :3], rtol=2e-6, atol=1e-7)


def test_row_sharding_bitwise(b200rt, gpu_ctx):
    """Tile sharding: rows rendered in 3 ragged bands are bitwise the full frame."""
    cam = b200rt.fixture_camera()
    params = b200rt.default_params(width=640, height=480)
    full, prim = gpu_ctx.render_whitted(cam, params)
    out = np.zeros_like(full)
    for (r0, rc) in [(0, 7), (7, 250), (257, 223)]:
        p = b200rt.copy_params(params, row_begin=r0, row_count=rc)
        gpu_ctx.render_whitted(cam, p, out_rgb=out, want_prim_id=False)
    assert np.array_equal(out.view(np.uint32), full.view(np.uint32))


def test_errors(b200rt, gpu_ctx):
    cam = b200rt.fixture_camera()
    with pytest.raises(b200rt.B200rtError) as e:
        gpu_ctx.render_whitted(cam, b200rt.default_params(depth=b200rt.MAX_DEPTH + 1))
    assert e.value.code == b200rt.ERR_UNSUPPORTED
    fresh = b200rt.Context(0)
    with pytest.raises(b200rt.B200rtError) as e:
        fresh.render_whitted(cam, b200rt.default_params(width=8, height=8))
    assert e.value.code == b200rt.ERR_NO_SCENE
    fresh.close()


# ---- stochastic tracer: wavefront == megakernel == brute-force exact cast ------------------------------------------
def test_grazing_rays_hit_at_infinity(b200rt, oracle, gpu_ctx, fixture_world):
    """Rays found by the 4K stochastic render whose reference n.dir is exactly 0 for one triangle while the fused dot
    product is not: the reference reports a hit at t = +inf (main.rs:204-231 accept inf / NaN).  The cast filter must
    keep such pairs whatever the sign of its own fused n.dir."""
    cases = [("3f959753 00000000 bf87819a", "bf00d52e 3f3c2e41 bee89aa4", 0, 37, 1, 30),
             ("40042e8a 4024e36a 400203d6", "befd3bb2 bead9492 bf4cdeb7", 0, -1, 0, 10),
             ("3f83e17c 00000000 bf0a1ae0", "beaf1a07 3f4ca2b9 befcf18c", 0, 37, 1, 31),
             ("bfd6e906 00000000 3fb82ed9", "3f7b6335 3e249b5d bdcb772a", 0, 36, 1, 25)]
    rays = np.zeros(len(cases) * 16, dtype=b200rt.RAY_DTYPE)
    for i in range(len(rays)):
        o, d, face, ex, exf, _ = cases[i % len(cases)]
        rays["origin"][i] = np.array([int(x, 16) for x in o.split()], dtype=np.uint32).view(np.float32)
        rays["direction"][i] = np.array([int(x, 16) for x in d.split()], dtype=np.uint32).view(np.float32)
        rays["face_direction"][i] = face; rays["exclude_prim"][i] = ex; rays["exclude_face"][i] = exf
    o = oracle.intersect(fixture_world.scene(), rays)
    for mode in (b200rt.CAST_TWO_PHASE, b200rt.CAST_BRUTE_EXACT):
        g = gpu_ctx.intersect(rays, mode)
        assert np.array_equal(g["prim_id"], o["prim_id"])
        assert np.array_equal(g["prim_id"][:4], [c[5] for c in cases])
        assert np.all(np.isinf(g["distance"]))


def test_tracers_and_cast_modes_bitwise(b200rt, gpu_ctx):
    """The wavefront tracer, the megakernel, and the megakernel with every pair through the exact test produce the same
    bits (one epoch in flight per pixel: 2560x1440 is above the wavefront's multi-epoch threshold)."""
    cam = b200rt.fixture_camera()
    out = {}
    for name, cm, tr in (("wave", b200rt.CAST_TWO_PHASE, b200rt.TRACER_WAVEFRONT),
                         ("mega", b200rt.CAST_TWO_PHASE, b200rt.TRACER_MEGAKERNEL),
                         ("brute", b200rt.CAST_BRUTE_EXACT, b200rt.TRACER_MEGAKERNEL)):
        p = b200rt.default_params(width=3840, height=2160, seed=11, tracer=tr, cast_mode=cm, row_begin=700, row_count=600)
        gpu_ctx.reset_stats()
        out[name] = (gpu_ctx.render_distributed(cam, p, 0, 3), gpu_ctx.stats())
    # the wavefront with one kernel pass per cast (the flow for scenes of > 4 lights) casts exactly the megakernel's rays
    os.environ["B200RT_WF_FUSED_LEVELS"] = "0"
    try:
        p = b200rt.default_params(width=3840, height=2160, seed=11, tracer=b200rt.TRACER_WAVEFRONT,
                                  cast_mode=b200rt.CAST_TWO_PHASE, row_begin=700, row_count=600)
        gpu_ctx.reset_stats()
        out["wave_unfused"] = (gpu_ctx.render_distributed(cam, p, 0, 3), gpu_ctx.stats())
    finally:
        del os.environ["B200RT_WF_FUSED_LEVELS"]
    for name in ("wave", "wave_unfused", "mega"):
        assert out[name][1]["samples"] == out["brute"][1]["samples"]
        assert np.array_equal(out[name][0].view(np.uint32), out["brute"][0].view(np.uint32)), name
    for name in ("wave_unfused", "mega"):
        assert out[name][1]["casts"] == out["brute"][1]["casts"]
    # fused levels: the get_shade that closes a sample after a missed bounce (main.rs:572-574) reuses the shadow rays
    # the hit already cast instead of casting them again
    assert 0.8 * out["brute"][1]["casts"] < out["wave"][1]["casts"] < out["brute"][1]["casts"]
    assert out["wave"][1]["wavefront_rounds"] > 0 and out["mega"][1]["wavefront_rounds"] == 0


@pytest.mark.parametrize("size", [(64, 48, 5), (333, 77, 3), (1920, 1080, 8)])
def test_wavefront_matches_megakernel(b200rt, gpu_ctx, size):
    """Small frames run several epochs in flight per pixel: same samples, sums differ only by fp32 summation order."""
    w, h, ep = size
    cam = b200rt.fixture_camera()
    a = gpu_ctx.render_distributed(cam, b200rt.default_params(width=w, height=h, seed=5, tracer=b200rt.TRACER_WAVEFRONT), 0, ep)
    m = gpu_ctx.render_distributed(cam, b200rt.default_params(width=w, height=h, seed=5, tracer=b200rt.TRACER_MEGAKERNEL), 0, ep)
    assert np.array_equal(a[..., 3], m[..., 3])
    np.testing.assert_allclose(a[..., :3], m[..., :3], rtol=2e-6, atol=1e-7)


def test_wavefront_depth0_and_many_lights(b200rt, oracle, gpu_ctx, fixture_world):
    """depth 0 (get_shade of the primary hit only, main.rs:525-527) and a scene with 6 lights (two chunks of shadow rays)."""
    cam = b200rt.fixture_camera()
    for depth in (0, 1):
        p = b200rt.default_params(width=200, height=150, seed=2, depth=depth)
        acc = gpu_ctx.render_distributed(cam, p, 0, 3)
        o_acc, _ = oracle.render_distributed(fixture_world.scene(), cam, p, 0, 3)
        assert np.array_equal(acc[..., 3], o_acc[..., 3])
        assert_stochastic_agreement(acc, o_acc, "fused levels / light counts", 3, min_psnr=35.0)
    w = b200rt.World.fixture()
    for k in range(3):
        w.push_light(b200rt.point_light([1.0 + k, 2.0, 1.5 - k], [0.3, 0.4, 0.5]))
    ctx = b200rt.Context(0)
    ctx.upload_scene(w)
    p = b200rt.default_params(width=160, height=120, seed=4)
    acc = ctx.render_distributed(cam, p, 0, 2)
    mega = ctx.render_distributed(cam, b200rt.copy_params(p, tracer=b200rt.TRACER_MEGAKERNEL), 0, 2)
    o_acc, _ = oracle.render_distributed(w.scene(), cam, p, 0, 2)
    assert np.array_equal(acc[..., 3], mega[..., 3])
    np.testing.assert_allclose(acc[..., :3], mega[..., :3], rtol=2e-6, atol=1e-7)
    # libm ulps (CUDA vs glibc powf / sincosf) can move a scattered ray across a silhouette: a sample or two may be
    # dropped by the is_normal filter on one side only
    assert (acc[..., 3] != o_acc[..., 3]).sum() <= 3
    assert_stochastic_agreement(acc, o_acc, "depth 0 / many lights", 2, min_psnr=35.0)
    ctx.close()


@pytest.mark.parametrize("extra", [1, 2])
def test_wavefront_fused_levels_light_counts(b200rt, gpu_ctx, extra):
    """Fused levels cast a hit's shadow rays and the next level's ray in one round: 4 lights = 4 shadow rays + the path
    ray = 5 work items per path (the capacity of the work list); 5 lights take the one-pass-per-cast flow (two chunks).
    Both must reproduce the megakernel: same samples, sums to fp32 summation order (several epochs in flight)."""
    cam = b200rt.fixture_camera()
    w = b200rt.World.fixture()
    w.push_light(b200rt.point_light([1.5, 2.5, 1.0], [0.4, 0.3, 0.5]))
    if extra == 2:
        w.push_light(b200rt.spot_light([0.5, 3.0, 1.5], [0.0, -1.0, -0.2], 0.9, 2.0, [0.6, 0.6, 0.4]))
    ctx = b200rt.Context(0)
    ctx.upload_scene(w)
    try:
        for (wd, ht, ep, depth) in ((200, 150, 3, 5), (96, 64, 2, 2), (64, 48, 2, 1)):
            p = b200rt.default_params(width=wd, height=ht, seed=6, depth=depth)
            ctx.reset_stats()
            acc = ctx.render_distributed(cam, p, 0, ep)
            st = ctx.stats()
            ctx.reset_stats()
            mega = ctx.render_distributed(cam, b200rt.copy_params(p, tracer=b200rt.TRACER_MEGAKERNEL), 0, ep)
            st_m = ctx.stats()
            assert np.array_equal(acc[..., 3], mega[..., 3])
            np.testing.assert_allclose(acc[..., :3], mega[..., :3], rtol=2e-6, atol=1e-7)
            assert st["samples"] == st_m["samples"]
            if extra == 1:
                assert st["casts"] <= st_m["casts"]          # fused: closing get_shade reuses the shadow rays
            else:
                assert st["casts"] == st_m["casts"]          # > 4 lights: the megakernel's casts exactly
    finally:
        ctx.close()


def test_obj_mesh_with_texcoords_and_normals_renders_like_the_oracle(b200rt, oracle, tmp_path):
    """N3 through the hot path: a two-model OBJ with `vt` and `vn` (a smooth-shaded dome over a textured ground plate)
    loaded with b200rt_world_load_obj_ex - interpolated vertex normals (main.rs:248-251) and interpolated uv feeding a
    procedural texture (main.rs:252, materials.rs:85-103) - renders the oracle's frame: identical hit ids, colours to 1e-4."""
    n = 12
    lines, vid = ["o dome"], {}
    for j in range(n + 1):
        for i in range(n + 1):
            u, v = i / n, j / n
            x, z = 2 * u - 1, 2 * v - 1
            y = 0.6 * (1 - x * x) * (1 - z * z)
            nx, nz = 1.2 * x * (1 - z * z), 1.2 * z * (1 - x * x)           # -dy/dx, -dy/dz
            nl = (nx * nx + 1 + nz * nz) ** 0.5
            lines += [f"v {x:.6f} {y:.6f} {z:.6f}", f"vt {u:.6f} {v:.6f}", f"vn {nx / nl:.6f} {1 / nl:.6f} {nz / nl:.6f}"]
            vid[(i, j)] = len(vid) + 1
    for j in range(n):
        for i in range(n):
            a, b_, c, d = vid[(i, j)], vid[(i + 1, j)], vid[(i + 1, j + 1)], vid[(i, j + 1)]
            lines.append(f"f {a}/{a}/{a} {d}/{d}/{d} {c}/{c}/{c} {b_}/{b_}/{b_}")   # a quad: fan of two, facing +y
    lines += ["g plate", "v -3 -0.01 -3", "v 3 -0.01 -3", "v 3 -0.01 3", "v -3 -0.01 3", "vt 0 0", "vt 1 0", "vt 1 1", "vt 0 1",
              "f -4/-4 -1/-1 -2/-2 -3/-3"]
    path = tmp_path / "dome.obj"
    path.write_text("\n".join(lines) + "\n")
    w = b200rt.World()
    dome = w.push_object(b200rt.generative_material(diffuse_fn=b200rt.DIFFUSE_CHECKER_UPV, freq=8.0, c0=(1.0, 0.2, 0.2), c1=(0.2, 0.2, 1.0),
                                                    shiness=0.3, smoothness=0.05))
    plate = w.push_object(b200rt.generative_material(diffuse_fn=b200rt.DIFFUSE_STRIPE_V, freq=12.0, c0=(0.9, 0.9, 0.9), c1=(0.3, 0.5, 0.3),
                                                     shiness=0.4, smoothness=0.3))
    ident = dict(scale_div=1.0, offset=(0, 0, 0), use_texcoords=True, use_normals=True)
    assert dome.load_obj_ex(str(path), model_index=0, **ident) == 2 * n * n
    assert plate.load_obj_ex(str(path), model_index=1, **ident) == 2
    w.push_light(b200rt.directional_light([-0.4, -1.0, -0.3], [1.0, 0.95, 0.9]))
    w.push_light(b200rt.point_light([1.5, 2.0, 1.5], [0.6, 0.6, 0.8]))
    cam = b200rt.fixture_camera()
    ctx = b200rt.Context(0)
    ctx.upload_scene(w)
    for cm in (b200rt.CAST_TWO_PHASE, b200rt.CAST_BVH):
        rgb, prim, _ = check_whitted(b200rt, oracle, ctx, w, b200rt.default_params(width=320, height=240, cast_mode=cm), cam)
    assert (prim >= 0).mean() > 0.3 and (prim < 2 * n * n).mean() > 0.05 and (prim >= 2 * n * n).any()
    # flat shading of the same file gives another frame: the normals of the file are really used
    w2 = b200rt.World()
    d2 = w2.push_object(b200rt.color_material(shiness=0.3, smoothness=0.05))
    d2.load_obj_ex(str(path), model_index=0, scale_div=1.0, offset=(0, 0, 0))
    w2.push_light(b200rt.directional_light([-0.4, -1.0, -0.3], [1.0, 0.95, 0.9]))
    ctx.upload_scene(w2)
    flat, _ = ctx.render_whitted(cam, b200rt.default_params(width=320, height=240))
    w3 = b200rt.World()
    d3 = w3.push_object(b200rt.color_material(shiness=0.3, smoothness=0.05))
    d3.load_obj_ex(str(path), model_index=0, scale_div=1.0, offset=(0, 0, 0), use_normals=True)
    w3.push_light(b200rt.directional_light([-0.4, -1.0, -0.3], [1.0, 0.95, 0.9]))
    ctx.upload_scene(w3)
    smooth, _ = ctx.render_whitted(cam, b200rt.default_params(width=320, height=240))
    assert np.abs(flat - smooth).max() > 0.02
    ctx.close()
