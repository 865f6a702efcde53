"""K2 on a multi-tile scene: b200rt_intersect_device on resident random rays against the fixture scene + an n x n
height-field mesh (tests/scene_util.py).   python tools/intersect_bench_mesh.py [n_grid=225] [n_rays=2^21]"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import __graft_entry__ as g
from scene_util import fixture_plus_mesh
b = g.load_package()
n_grid = int(sys.argv[1]) if len(sys.argv) > 1 else 225
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 21
world, ntri = fixture_plus_mesh(b, tempfile.mkdtemp(), n_grid)
ctx = b.Context(0)
ctx.upload_scene(world)
sc = world.scene()
rng = np.random.default_rng(0)
rays = np.zeros(n, dtype=b.RAY_DTYPE)
o = rng.uniform(-2.5, 2.5, size=(n, 3)).astype(np.float32); o[:, 1] = rng.uniform(-0.5, 3.0, size=n)
d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
rays["origin"] = o; rays["direction"] = d.astype(np.float32); rays["face_direction"] = rng.integers(0, 3, size=n); rays["exclude_prim"] = -1
d_rays = torch.from_numpy(rays.view(np.uint8)).cuda()
d_hits = torch.empty(n * b.HIT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
peak = ctx.device_info()["sm_count"] * 128 * 2 * 1.965e9
flop = (36.0 * sc.n_triangles + 28.0 * sc.n_spheres) * n
for rep in range(3):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    ctx.intersect_device(d_rays.data_ptr(), n, d_hits.data_ptr(), b.CAST_TWO_PHASE, torch.cuda.current_stream().cuda_stream)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
print(os.environ.get("B200RT_INTERSECT", "rl"), f"{sc.n_triangles} triangles, {n} rays: {ms:.1f} ms  {n * sc.n_triangles / ms / 1e6:.1f} Gpairs/s  {100 * flop / (ms * 1e-3) / peak:.1f} % of the FP32 roofline")
