"""Small workloads that touch every kernel family once and cross-check them against each other (cast modes, tracers, the
device-side round loop against host-enqueued rounds, one-tile and tiled scenes, the image finishers): a quick whole-library
sanity run on a GPU, and the workload for `compute-sanitizer --tool memcheck python tools/sanitize_run.py` where the tool is
available."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import __graft_entry__ as g
from scene_util import fixture_plus_mesh
b = g.load_package()
cam = b.fixture_camera()
ctx = b.Context(0)
ctx.upload_scene(b.World.fixture())
p = b.default_params(width=96, height=64, seed=3)
for cm in (b.CAST_TWO_PHASE, b.CAST_BVH, b.CAST_BRUTE_EXACT):
    rgb, prim = ctx.render_whitted(cam, b.copy_params(p, cast_mode=cm))
ref = ctx.render_distributed(cam, p, 0, 3)                                             # wavefront, device-side round loop
os.environ["B200RT_WF_GRAPH"] = "0"
host = ctx.render_distributed(cam, p, 0, 3)                                            # ... host-enqueued rounds
del os.environ["B200RT_WF_GRAPH"]
assert np.array_equal(ref.view(np.uint32), host.view(np.uint32))
for tr, cm in ((b.TRACER_MEGAKERNEL, b.CAST_TWO_PHASE), (b.TRACER_WAVEFRONT, b.CAST_BVH), (b.TRACER_MEGAKERNEL, b.CAST_BVH)):
    acc = ctx.render_distributed(cam, b.copy_params(p, tracer=tr, cast_mode=cm), 0, 3)
    assert np.array_equal(acc[..., 3], ref[..., 3])
rng = np.random.default_rng(0)
rays = np.zeros(5000, dtype=b.RAY_DTYPE)
rays["origin"] = rng.uniform(-2, 2, size=(5000, 3)); d = rng.normal(size=(5000, 3)); rays["direction"] = d / np.linalg.norm(d, axis=1, keepdims=True)
rays["exclude_prim"] = -1; rays["face_direction"] = rng.integers(0, 3, size=5000)
rays["origin"][:50, 0] = np.nan; rays["origin"][50:100] *= 1e6
h0 = ctx.intersect(rays, b.CAST_TWO_PHASE); h1 = ctx.intersect(rays, b.CAST_BVH)
assert np.array_equal(h0["prim_id"], h1["prim_id"])
ctx.close()
world, ntri = fixture_plus_mesh(b, tempfile.mkdtemp(), n=17)                             # 512 + 64 triangles: the tiled TMA cast
ctx = b.Context(0)
ctx.upload_scene(world)
a0 = ctx.render_distributed(cam, p, 0, 2)
a1 = ctx.render_distributed(cam, b.copy_params(p, cast_mode=b.CAST_BVH), 0, 2)
assert np.array_equal(a0.view(np.uint32), a1.view(np.uint32))
h0 = ctx.intersect(rays, b.CAST_TWO_PHASE); h1 = ctx.intersect(rays, b.CAST_BVH)
assert np.array_equal(h0["prim_id"], h1["prim_id"])
img = np.ascontiguousarray(np.clip(ctx.render_whitted(cam, p)[0], 0, 4).astype(np.float32))
ctx.post_process(img); ctx.encode_srgb8(img)
ctx.close()
print("sanitize_run ok")
