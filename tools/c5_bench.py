"""C5: fixture scene + the 100 352-triangle synthetic OBJ height field, 4000x2500 = 10 M "photons" (one
stochastic sample per pixel, 1 epoch, depth 5) accumulated with PhotonAccumulator semantics, tile-sharded in 8 bands
of rows.  Times ONE band (what one of the 8 GPUs renders) on this GPU and reports pair tests / s against the FP32
roofline.   python tools/c5_bench.py [band 0..7] [n_grid]"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import __graft_entry__ as g
from scene_util import fixture_plus_mesh
b = g.load_package()
band = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_grid = int(sys.argv[2]) if len(sys.argv) > 2 else 225
tmp = tempfile.mkdtemp()
world, ntri = fixture_plus_mesh(b, tmp, n_grid)
ctx = b.Context(0)
t0 = time.perf_counter(); ctx.upload_scene(world); t_up = time.perf_counter() - t0
W, H = 4000, 2500
r0, r1 = band * H // 8, (band + 1) * H // 8
p = b.default_params(width=W, height=H, seed=0, row_begin=r0, row_count=r1 - r0)
cam = b.fixture_camera()
ctx.render_distributed(cam, p, 0, 1)          # warm-up
ctx.reset_stats()
acc = ctx.render_distributed(cam, p, 0, 1)
s = ctx.stats()
flops = s["tri_pair_tests"] * 36.0 + s["sph_pair_tests"] * 28.0
peak = ctx.device_info()["sm_count"] * 128 * 2 * 1.965e9
print(f"C5 band {band} rows [{r0},{r1}) of {W}x{H}, {ntri} mesh triangles (+64 +4 spheres), upload {t_up*1e3:.0f} ms")
print(f"  samples {s['samples']} casts {s['casts']} pair tests {s['tri_pair_tests'] + s['sph_pair_tests']:.3e} kernel {s['kernel_ms']:.1f} ms")
print(f"  {s['samples'] / s['kernel_ms'] / 1e3:.2f} Mrays/s (one GPU, one band)  {(s['tri_pair_tests'] + s['sph_pair_tests']) / s['kernel_ms'] / 1e6:.1f} Gpairs/s  "
      f"{flops / (s['kernel_ms'] * 1e-3) / 1e12:.2f} TFLOP/s algorithmic = {100 * flops / (s['kernel_ms'] * 1e-3) / peak:.1f} % of the FP32 roofline; "
      f"exact tests / cast {s['exact_confirms'] / max(s['casts'], 1):.2f}")
if os.environ.get("B200RT_LIB", "").endswith("diag.so"):
    fb = s['certify_fallbacks']
    print(f"  untrusted finite rays: origin beyond the bound {(fb >> 40) & 0xfff}, |dir|^2 off {fb >> 52}, of {s['casts']} casts; certify fallbacks {fb & ((1 << 40) - 1)}")
