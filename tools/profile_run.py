"""Small fixed workload for ncu: one distributed launch (960x540, 4 epochs) and one Whitted launch (960x540)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
b = g.load_package()
ctx = b.Context(0)
ctx.upload_scene(b.World.fixture())
cam = b.fixture_camera()
p = b.default_params(width=960, height=540, seed=0)
mode = sys.argv[1] if len(sys.argv) > 1 else "both"
if mode in ("both", "distributed"):
    acc = ctx.render_distributed(cam, p, 0, 4)
    print("distributed", ctx.stats())
if mode in ("both", "whitted"):
    ctx.reset_stats()
    rgb, prim = ctx.render_whitted(cam, p)
    print("whitted", ctx.stats())
if mode in ("both", "intersect"):
    rng = np.random.default_rng(0)
    n = 1 << 22
    rays = np.zeros(n, dtype=b.RAY_DTYPE)
    rays["origin"] = rng.uniform(-2, 2, size=(n, 3)).astype(np.float32) + np.array([0, 1.5, 0], dtype=np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["direction"] = d; rays["exclude_prim"] = -1
    ctx.reset_stats()
    hits = ctx.intersect(rays)
    print("intersect", ctx.stats())
