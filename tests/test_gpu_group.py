"""Device groups of the C ABI (include/b200rt.h, b200rt_group_*): one frame over several GPUs with NCCL inside the library.

A group of one GPU needs no NCCL and must reproduce the single-GPU entry points bit for bit; with two or more GPUs in
the box (gpurun --gpus 2) the single-process group (ncclCommInitAll, one host thread per device) shards epochs / rows
and reduces / gathers on rank 0: rows are bitwise the single-GPU frame, epochs equal it up to the fp32 summation order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return torch.cuda.device_count()


def test_group_of_one_equals_the_context(b200rt, gpu_ctx, fixture_world):
    cam = b200rt.fixture_camera()
    g = b200rt.Group([0])
    assert g.size() == (1, 1)
    g.upload_scene(fixture_world)
    p = b200rt.default_params(width=320, height=200, seed=3)
    acc = g.render_distributed(cam, p, 2, 3)
    ref = gpu_ctx.render_distributed(cam, p, 2, 3)
    assert np.array_equal(acc.view(np.uint32), ref.view(np.uint32))
    rows = g.render_distributed(cam, p, 0, 1, by_rows=True)
    assert np.array_equal(rows.view(np.uint32), gpu_ctx.render_distributed(cam, p, 0, 1).view(np.uint32))
    rgb, prim = g.render_whitted(cam, p)
    r_rgb, r_prim = gpu_ctx.render_whitted(cam, p)
    assert np.array_equal(rgb.view(np.uint32), r_rgb.view(np.uint32)) and np.array_equal(prim, r_prim)
    band = b200rt.copy_params(p, row_begin=50, row_count=40)
    rgb_b, _ = g.render_whitted(cam, band)
    assert np.array_equal(rgb_b[50:90].view(np.uint32), r_rgb[50:90].view(np.uint32)) and not rgb_b[:50].any() and not rgb_b[90:].any()
    assert g.last_render_ms() > 0.0
    st = g.member(0).stats()
    assert st["casts"] > 0
    g.close()


def test_group_rejects_bad_arguments(b200rt):
    lib = b200rt.load_library()
    import ctypes as C
    h = C.c_void_p()
    assert lib.b200rt_group_create(None, 1, C.byref(h)) == b200rt.ERR_INVALID
    ids = (C.c_int * 2)(0, 0)
    assert lib.b200rt_group_create(ids, 2, C.byref(h)) == b200rt.ERR_INVALID          # the same device twice
    ids = (C.c_int * 1)(9999)
    assert lib.b200rt_group_create(ids, 1, C.byref(h)) == b200rt.ERR_NO_DEVICE
    assert lib.b200rt_group_create_rank(0, 3, 2, None, 0, C.byref(h)) == b200rt.ERR_INVALID
    assert lib.b200rt_group_destroy(None) == b200rt.ERR_INVALID


def test_intersect_rejects_values_the_reference_types_cannot_hold(b200rt, gpu_ctx):
    rays = np.zeros(4, dtype=b200rt.RAY_DTYPE)
    rays["direction"] = (0.0, 0.0, 1.0)
    rays["exclude_prim"] = -1
    gpu_ctx.intersect(rays)
    for field, value in (("face_direction", 3), ("exclude_face", 7), ("exclude_prim", -2)):
        bad = rays.copy()
        bad[field][2] = value
        with pytest.raises(b200rt.B200rtError) as e:
            gpu_ctx.intersect(bad)
        assert e.value.code == b200rt.ERR_INVALID
    ok = rays.copy()
    ok["exclude_prim"][1] = 1 << 20                 # beyond the scene's primitives: legal, never matches
    gpu_ctx.intersect(ok)


@pytest.mark.skipif("n_gpus() < 2")
def test_two_gpu_group_matches_one_gpu(b200rt, gpu_ctx, fixture_world):
    cam = b200rt.fixture_camera()
    n = min(n_gpus(), 4)
    g = b200rt.Group(list(range(n)))
    assert g.size() == (n, n)
    g.upload_scene(fixture_world)
    p = b200rt.default_params(width=640, height=360, seed=5)
    acc = g.render_distributed(cam, p, 0, 8)                        # epochs sharded, ncclReduce(sum) to rank 0
    ref = gpu_ctx.render_distributed(cam, p, 0, 8)
    assert np.array_equal(acc[..., 3], ref[..., 3])
    np.testing.assert_allclose(acc[..., :3], ref[..., :3], rtol=2e-6, atol=1e-7)
    rows = g.render_distributed(cam, p, 0, 2, by_rows=True)         # rows sharded, gathered on rank 0
    assert np.array_equal(rows.view(np.uint32), gpu_ctx.render_distributed(cam, p, 0, 2).view(np.uint32))
    rgb, prim = g.render_whitted(cam, p)
    r_rgb, r_prim = gpu_ctx.render_whitted(cam, p)
    assert np.array_equal(rgb.view(np.uint32), r_rgb.view(np.uint32)) and np.array_equal(prim, r_prim)
    tiny = b200rt.default_params(width=64, height=n - 1 if n > 1 else 1, seed=1)     # fewer rows than ranks: a rank renders nothing
    t_rgb, _ = g.render_whitted(cam, tiny)
    assert np.array_equal(t_rgb.view(np.uint32), gpu_ctx.render_whitted(cam, tiny)[0].view(np.uint32))
    g.close()


def test_row_strips_assemble_the_frame_bitwise(b200rt, gpu_ctx):
    """b200rt_render_distributed_strips_device: the strips s % 3 == part of three calls fill disjoint rows, and their sum is
    bitwise the whole-frame render (what a 3-rank row-sharded group reduces on rank 0); ragged last strip, band of rows."""
    import ctypes as C
    import torch
    lib = b200rt.load_library()
    cam = b200rt.fixture_camera()
    for (w, h, band) in ((200, 77, None), (160, 120, (13, 90))):
        p = b200rt.default_params(width=w, height=h, seed=9)
        if band:
            p = b200rt.copy_params(p, row_begin=band[0], row_count=band[1])
        ref = gpu_ctx.render_distributed(cam, p, 1, 2)
        total = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
        seen = torch.zeros((h, w), dtype=torch.int32, device="cuda")
        for part in range(3):
            d = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
            rc = lib.b200rt_render_distributed_strips_device(gpu_ctx._h, C.byref(cam), C.byref(p), 1, 2, C.c_void_p(d.data_ptr()),
                                                             C.c_void_p(torch.cuda.current_stream().cuda_stream), 16, 3, part)
            assert rc == 0, rc
            torch.cuda.synchronize()
            rows = (d[..., 3] != 0).any(dim=1)
            seen += (d != 0).any(dim=2).to(torch.int32)
            r0 = band[0] if band else 0
            idx = torch.nonzero(rows).flatten().cpu().numpy()
            assert all(((int(r) - r0) // 16) % 3 == part for r in idx)          # only its own strips
            total += d
        assert int(seen.max()) <= 1
        assert np.array_equal(total.cpu().numpy().view(np.uint32), ref.view(np.uint32))
    bad = b200rt.default_params(width=64, height=64, tracer=b200rt.TRACER_MEGAKERNEL)
    d = torch.zeros((64, 64, 4), dtype=torch.float32, device="cuda")
    assert lib.b200rt_render_distributed_strips_device(gpu_ctx._h, C.byref(cam), C.byref(bad), 0, 1, C.c_void_p(d.data_ptr()), None, 16, 2, 0) == b200rt.ERR_UNSUPPORTED
    assert lib.b200rt_render_distributed_strips_device(gpu_ctx._h, C.byref(cam), C.byref(p), 0, 1, C.c_void_p(d.data_ptr()), None, 16, 2, 2) == b200rt.ERR_INVALID
