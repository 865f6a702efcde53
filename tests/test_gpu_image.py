"""GPU parity of the image finishers: post_process (main.rs:748-762) and the sRGB u8 encode
(image.rs:55-66) on the device against the oracle, and the whole GPU pipeline against the reference's own PNG."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_post_process_random_images_bit_exact(b200rt, oracle, gpu_ctx):
    """The radix select finds exactly the element the reference's sort + index picks; the division is IEEE."""
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 255, 256, 257, 1000, 64 * 48, 333 * 77):
        img = rng.gamma(0.7, 0.6, size=(n, 3)).astype(np.float32)
        g, gp = gpu_ctx.post_process(img)
        o, op = oracle.post_process(img)
        assert np.float32(gp).view(np.uint32) == np.float32(op).view(np.uint32), (n, gp, op)
        assert np.array_equal(bits(g), bits(o)), n
    # many equal lumas (ties inside one radix bucket), a constant image
    img = np.repeat(rng.integers(1, 5, size=(5000, 1)).astype(np.float32) * 0.25, 3, axis=1)
    g, gp = gpu_ctx.post_process(img); o, op = oracle.post_process(img)
    assert gp == op and np.array_equal(bits(g), bits(o))
    img = np.full((4096, 3), 0.3, dtype=np.float32)
    g, gp = gpu_ctx.post_process(img); o, op = oracle.post_process(img)
    assert gp == op and np.array_equal(bits(g), bits(o))


def test_post_process_edge_cases(b200rt, oracle, gpu_ctx):
    """Zero, subnormal, NaN, inf and negative lumas: only NORMAL lumas are ranked (main.rs:751); an image without any
    (the reference would panic on the index) and a p98 below f32::EPSILON are left untouched."""
    rng = np.random.default_rng(4)
    img = rng.normal(0.2, 0.5, size=(20000, 3)).astype(np.float32)        # negative lumas take part, in numeric order
    img[::7] = 0.0
    img[3::11] = np.float32(1e-41)                                        # subnormal
    img[5::13, 1] = np.nan
    img[8::17, 2] = np.inf
    g, gp = gpu_ctx.post_process(img)
    o, op = oracle.post_process(img)
    assert np.float32(gp).view(np.uint32) == np.float32(op).view(np.uint32)
    assert np.array_equal(np.isnan(g), np.isnan(o))
    m = ~np.isnan(o)
    assert np.array_equal(bits(g)[m], bits(o)[m])
    for img in (np.zeros((300, 3), np.float32), np.full((300, 3), 1e-9, np.float32), np.zeros((0, 3), np.float32)):
        g, gp = gpu_ctx.post_process(img)
        o, op = oracle.post_process(img) if len(img) else (img, 0.0)
        assert gp == op == 0.0 and np.array_equal(bits(g), bits(img))


def test_encode_srgb8_matches_oracle(b200rt, oracle, gpu_ctx):
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-0.2, 1.3, size=300000), np.linspace(0, 0.0031308 * 2, 4097),
                        [np.nan, np.inf, -np.inf, 0.0, 1.0, 0.0031308]]).astype(np.float32)
    x = x[: (len(x) // 3) * 3].reshape(-1, 3)
    g = gpu_ctx.encode_srgb8(x).astype(int)
    o = oracle.encode_srgb8(x).astype(int)
    d = np.abs(g - o)
    # CUDA powf vs glibc powf: a value within an ulp of a rounding boundary may land on the neighbouring code
    assert d.max() <= 1 and (d > 0).mean() < 1e-4, (d.max(), (d > 0).mean())
    lin = (x <= 0.0031308) | ~np.isfinite(x)
    assert np.array_equal(g[lin], o[lin])                                 # the linear segment and the clamps are exact


def test_gpu_pipeline_reproduces_reference_png(b200rt, oracle, gpu_ctx):
    """render_whitted -> post_process -> encode, all on the GPU, against report/out_single_epoch.png (the committed
    probes): the same bar the oracle is pinned with (tests/test_oracle_golden.py)."""
    gold = json.load(open(os.path.join(GOLD, "out_single_epoch_probe.json")))
    rgb, prim = gpu_ctx.render_whitted(b200rt.fixture_camera(), b200rt.default_params())
    assert rgb.shape == (gold["height"], gold["width"], 3)
    pp, p98 = gpu_ctx.post_process(rgb)
    o_pp, o_p98 = oracle.post_process(rgb)
    assert p98 == o_p98 and np.array_equal(bits(pp), bits(o_pp))          # exact on the GPU's own frame
    img = gpu_ctx.encode_srgb8(pp).astype(int)
    probes = np.array(gold["probes_y_x_r_g_b"])
    diff = np.abs(img[probes[:, 0], probes[:, 1]] - probes[:, 2:5]).max(axis=1)
    assert (diff <= 2).mean() >= 0.998 and np.median(diff) <= 1
    assert abs(int((img.max(axis=2) == 0).sum()) - gold["black_pixels"]) <= 0.02 * gold["black_pixels"]
    np.testing.assert_allclose(img.reshape(-1, 3).mean(axis=0), gold["mean_rgb"], atol=1.0)


def test_device_resident_finish_of_a_stochastic_frame(b200rt, oracle, gpu_ctx):
    """accumulators -> resolve -> post_process -> encode without leaving HBM (torch tensors own the buffers)."""
    torch = pytest.importorskip("torch")
    cam = b200rt.fixture_camera()
    p = b200rt.default_params(width=320, height=240, seed=9)
    acc = torch.zeros((240, 320, 4), dtype=torch.float32, device="cuda")
    rgb = torch.empty((240, 320, 3), dtype=torch.float32, device="cuda")
    u8 = torch.empty((240, 320, 3), dtype=torch.uint8, device="cuda")
    p98 = torch.zeros(1, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    gpu_ctx.render_distributed_device(cam, p, 0, 8, acc.data_ptr(), st)
    gpu_ctx.resolve_device(acc.data_ptr(), rgb.data_ptr(), 320 * 240, st)
    mean = rgb.cpu().numpy().copy()
    gpu_ctx.post_process_device(rgb.data_ptr(), 320 * 240, p98.data_ptr(), st)
    gpu_ctx.encode_srgb8_device(rgb.data_ptr(), 320 * 240 * 3, u8.data_ptr(), st)
    torch.cuda.synchronize()
    o_pp, o_p98 = oracle.post_process(mean)
    assert float(p98.item()) == o_p98 and np.array_equal(bits(rgb.cpu().numpy()), bits(o_pp))
    d = np.abs(u8.cpu().numpy().astype(int) - oracle.encode_srgb8(o_pp).astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3


def test_render_main_progressive_renormalisation(b200rt, oracle, gpu_ctx, fixture_world, tmp_path):
    """The whole render part of main() (main.rs:1086-1173): Whitted frame, then epochs added into the image that is
    re-normalised by its p99 luma after every frame.  GPU chain (render, post_process, encode, PNG) against the same
    chain on the oracle; the written PNG decodes to the last frame."""
    from PIL import Image
    cam = b200rt.fixture_camera()
    p = b200rt.default_params(width=160, height=120, seed=9)
    frames = {}
    out = tmp_path / "out.png"
    img = b200rt.render_main(gpu_ctx, cam, p, 3, out_path=str(out), on_frame=lambda k, u8: frames.__setitem__(k, u8.copy()))
    # oracle chain
    o_img = np.zeros((120, 160, 3), dtype=np.float32)
    rgb, _, _ = oracle.render_whitted(fixture_world.scene(), cam, p)
    o_img, _ = oracle.post_process(o_img + rgb)
    o_frames = {0: oracle.encode_srgb8(o_img)}
    for i in range(3):
        acc, _ = oracle.render_distributed(fixture_world.scene(), cam, p, i, 1)
        o_img, _ = oracle.post_process(o_img + acc[..., :3])
        o_frames[i + 1] = oracle.encode_srgb8(o_img)
    assert sorted(frames) == [0, 1, 2, 3]
    for k in frames:
        d = np.abs(frames[k].astype(np.int32) - o_frames[k].astype(np.int32))
        # libm ulps move a few values across a code boundary (<= 1 level); a handful of stochastic samples land on
        # another primitive at silhouettes (see test_distributed_samples_match_oracle)
        assert (d > 1).mean() < 2e-3, (k, (d > 1).mean())
    fin = np.isfinite(o_img).all(axis=2)
    err = np.abs(img - o_img)[fin] / np.maximum(np.abs(o_img)[fin], 1e-3)
    assert (err > 1e-3).mean() < 2e-3
    assert np.array_equal(np.asarray(Image.open(out).convert("RGB")), frames[3])



def test_color_pow_error_bound(b200rt, gpu_ctx):
    """color_pow (csrc/rt_math.cuh) - the power behind the Phong lobe (materials.rs:63), the spot cone (lights.rs:62-64) and
    the opaque decay (main.rs:508 / 605), whose results are only ever colours - against f64 pow: absolute error <= 2.5e-7,
    relative error <= 3e-7 max(1, |e log2 x|) over every exponent of the fixture scene and the extremes the builder
    accepts; subnormal results are kept (the is_normal filter sees what the reference sees), and the arguments outside
    its fast path fall through to powf."""
    import ctypes as C
    lib = b200rt.load_library()
    rng = np.random.default_rng(5)
    f32p = C.POINTER(C.c_float)

    def dev(x, e):
        x = np.ascontiguousarray(x, dtype=np.float32); e = np.ascontiguousarray(np.broadcast_to(np.float32(e), x.shape), dtype=np.float32)
        out = np.zeros_like(x)
        rc = lib.b200rt_dev_color_pow(gpu_ctx._h, x.ctypes.data_as(f32p), e.ctypes.data_as(f32p), out.ctypes.data_as(f32p), x.size)
        assert rc == b200rt.OK
        return out

    eps = np.finfo(np.float32).eps
    exps = [1 / (s + eps) for s in (1.0, 0.7, 0.2, 0.01, 0.001, 1e-5)] + [1.0 + eps, 0.5, 3.0, 2.5e-3, 8.4e6, 37.5]
    worst_abs = worst_rel = 0.0
    for e in exps:
        e = np.float32(e)
        x = np.concatenate([rng.random(1 << 16), 1.0 - rng.random(1 << 16) * min(1.0, 60.0 / float(e)),
                            [1.0, np.nextafter(np.float32(1), np.float32(2)), 0.5, 0.25, 1e-20, 1e-37, 3e-38]]).astype(np.float32)
        got = dev(x, e).astype(np.float64)
        want = np.power(x.astype(np.float64), float(e))
        y = np.abs(float(e) * np.log2(x.astype(np.float64)))
        big = want >= 2.0 ** -126
        worst_abs = max(worst_abs, float(np.abs(got - want).max()))
        rel = np.abs(got[big] - want[big]) / want[big] / np.maximum(1.0, y[big])
        worst_rel = max(worst_rel, float(rel.max()))
        sub = (want < 2.0 ** -126) & (want >= 2.0 ** -149)
        if sub.any():    # subnormal results survive (MUFU.EX2 alone would flush them)
            assert (got[sub] > 0).all() and np.abs(got[sub] - want[sub]).max() <= 2.0 ** -140, float(e)
        assert (got[want < 2.0 ** -151] == 0).all()
    print(f"color_pow: max abs err {worst_abs:.2e}, max rel err / max(1, |e log2 x|) {worst_rel:.2e}")
    assert worst_abs <= 2.5e-7 and worst_rel <= 3e-7
    # outside the fast path: the libm answers
    assert dev([0.0, 0.0, 0.5, 2.0, np.inf], [2.0, 0.0, 0.0, 0.0, 2.0]).tolist() == [0.0, 1.0, 1.0, 1.0, np.inf]
    assert np.isnan(dev([np.nan], 2.0))[0] and np.isnan(dev([-0.5], 0.5))[0]
    assert dev([0.25], 0.5)[0] == 0.5 and dev([1.0], 1e5)[0] == 1.0
