"""C5 through the acceleration structure (B200RT_CAST_BVH): the whole 4000x2500 one-sample frame
of the fixture scene + the 100 352-triangle height field on this GPU, against the brute-force two-phase cast on a band of
rows (bitwise the same accumulators).   python tools/c5_bvh_bench.py [n_grid]"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import __graft_entry__ as g
from scene_util import fixture_plus_mesh
b = g.load_package()
n_grid = int(sys.argv[1]) if len(sys.argv) > 1 else 225
world, ntri = fixture_plus_mesh(b, tempfile.mkdtemp(), n_grid)
ctx = b.Context(0)
t0 = time.perf_counter(); ctx.upload_scene(world); t_up = time.perf_counter() - t0
W, H = 4000, 2500
cam = b.fixture_camera()
pb = b.default_params(width=W, height=H, seed=0, cast_mode=b.CAST_BVH)
ctx.render_distributed(cam, b.copy_params(pb, row_begin=1000, row_count=64), 0, 1)     # warm-up
ctx.reset_stats()
t0 = time.perf_counter(); acc = ctx.render_distributed(cam, pb, 0, 1); wall = time.perf_counter() - t0
s = ctx.stats()
print(f"C5 {W}x{H}, {ntri} mesh triangles (+64 +4 spheres), upload (records + trees) {t_up*1e3:.0f} ms")
print(f"  BVH: whole frame kernel {s['kernel_ms']:.1f} ms (wall {wall*1e3:.0f} ms) = {s['samples_generated'] / s['kernel_ms'] / 1e3 if 'samples_generated' in s else W*H / s['kernel_ms'] / 1e3:.1f} Mrays/s; "
      f"casts {s['casts']}, exact tests / cast {s['exact_confirms'] / max(s['casts'], 1):.1f}, ordered walks {s['certify_fallbacks']}, rounds {s['wavefront_rounds']}")
# the same rows through the brute-force two-phase cast: bitwise
r0, rn = 1100, 120
pt = b.default_params(width=W, height=H, seed=0, row_begin=r0, row_count=rn)
ctx.reset_stats()
ref = ctx.render_distributed(cam, pt, 0, 1)
s2 = ctx.stats()
same = np.array_equal(ref[r0:r0 + rn].view(np.uint32), acc[r0:r0 + rn].view(np.uint32))
print(f"  two-phase brute force, rows [{r0},{r0 + rn}): kernel {s2['kernel_ms']:.1f} ms ({s2['kernel_ms'] * H / rn:.0f} ms per frame at this rate); "
      f"accumulators bitwise equal to the BVH frame's rows: {same}")
