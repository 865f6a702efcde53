"""K2 (World::cast on its own) throughput on B200: b200rt_intersect_device on resident rays, CUDA events.

Two ray populations of the fixture scene: uniformly random rays (all miss / hit mixes) and the primary camera
rays of a 2560x1440 frame (coherent, the population the tracers produce).  Prints casts/s and the algorithmic
FP32 roofline fraction (36 flop per ray x triangle, 28 per ray x sphere)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as g
b = g.load_package()
ctx = b.Context(0)
world = b.World.fixture()
ctx.upload_scene(world)
sc = world.scene()
info = ctx.device_info()
peak = info["sm_count"] * 128 * 2 * 1.965e9
flop_per_cast = 36.0 * sc.n_triangles + 28.0 * sc.n_spheres


def camera_rays(w, h):
    cam = b.fixture_camera()
    toward = np.array(cam.toward, dtype=np.float64); toward /= np.linalg.norm(toward)
    up = np.array(cam.up, dtype=np.float64)
    right = np.cross(toward, up); right /= np.linalg.norm(right)
    up2 = np.cross(right, toward); up2 /= np.linalg.norm(up2)
    t = np.tan(cam.fovy / 2)
    ys, xs = np.mgrid[0:h, 0:w]
    cy = (h / 2 - ys) / h; cx = (xs - w / 2) / h
    d = cx[..., None] * (t * right) + cy[..., None] * (t * up2) + toward
    d /= np.linalg.norm(d, axis=2, keepdims=True)
    rays = np.zeros(w * h, dtype=b.RAY_DTYPE)
    rays["origin"] = (np.array(cam.center) + toward * cam.near).astype(np.float32)
    rays["direction"] = d.reshape(-1, 3).astype(np.float32)
    rays["exclude_prim"] = -1
    return rays


def random_rays(n):
    rng = np.random.default_rng(0)
    rays = np.zeros(n, dtype=b.RAY_DTYPE)
    rays["origin"] = rng.uniform(-2, 2, size=(n, 3)).astype(np.float32) + np.array([0, 1.5, 0], dtype=np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["direction"] = d; rays["exclude_prim"] = -1
    return rays


out = {}
for name, rays in (("random_4Mi", random_rays(1 << 22)), ("camera_2560x1440", camera_rays(2560, 1440))):
    n = len(rays)
    d_rays = torch.from_numpy(rays.view(np.uint8)).cuda()
    d_hits = torch.empty(n * b.HIT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for mode, mname in ((b.CAST_TWO_PHASE, "two_phase"),):
        for _ in range(3):
            ctx.intersect_device(d_rays.data_ptr(), n, d_hits.data_ptr(), mode, st)
        torch.cuda.synchronize()
        ctx.reset_stats()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a, c in ev:
            a.record(); ctx.intersect_device(d_rays.data_ptr(), n, d_hits.data_ptr(), mode, st); c.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(c) for a, c in ev)
        med = ms[len(ms) // 2]
        s = ctx.stats()
        res = {"rays": n, "ms_median": med, "ms_min": ms[0], "Gcasts_per_s": n / med / 1e6,
               "algorithmic_TFLOPs": n * flop_per_cast / (med * 1e-3) / 1e12,
               "roofline_frac": n * flop_per_cast / (med * 1e-3) / peak,
               "exact_tests_per_cast": s["exact_confirms"] / max(s["casts"], 1),
               "fallbacks_per_cast": s["certify_fallbacks"] / max(s["casts"], 1)}
        out[f"{name}/{mname}"] = res
        print(name, mname, json.dumps(res), flush=True)
