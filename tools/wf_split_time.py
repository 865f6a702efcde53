import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
b = g.load_package()
ctx = b.Context(0)
ctx.upload_scene(b.World.fixture())
cam = b.fixture_camera()
w, h, e = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "3840x2160x16").split("x"))
p = b.default_params(width=w, height=h, seed=0)
ctx.render_distributed(cam, p, 0, e)
ctx.set_kernel_timing(True)
ctx.reset_stats()
ctx.render_distributed(cam, p, 0, e)
s = ctx.stats()
prim = w * h * e if s.get("primary_kernel_ms", 0.0) > 0.0 else 0      # round 0 runs in its own kernel: not the dominant one
sc_ = b.World.fixture().scene()
flops = (s["tri_pair_tests"] - prim * sc_.n_triangles) * 36.0 + (s["sph_pair_tests"] - prim * sc_.n_spheres) * 28.0
peak = 148 * 128 * 2 * 1.965e9
print(os.environ.get("B200RT_LIB", "default"), os.environ.get("B200RT_WF_CAST", "default"),
      f"total {s['kernel_ms']:.1f} cast {s['cast_kernel_ms']:.1f} primary {s.get('primary_kernel_ms', 0.0):.1f} logic {s['logic_kernel_ms']:.1f}",
      f"| cast roofline {flops / (s['cast_kernel_ms'] * 1e-3) / peak:.3f}",
      f"| exact/cast {s['exact_confirms'] / max(s['casts'], 1):.3f} fallback/cast {s['certify_fallbacks'] / max(s['casts'], 1):.5f} casts {s['casts']}", flush=True)
