"""Small fixed wavefront workload for ncu: fixture scene, WxH, E epochs (default 1920x1080 x 2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
b = g.load_package()
ctx = b.Context(0)
ctx.upload_scene(b.World.fixture())
cam = b.fixture_camera()
w, h, e = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1920x1080x2").split("x"))
p = b.default_params(width=w, height=h, seed=0, tracer=b.TRACER_WAVEFRONT)
acc = ctx.render_distributed(cam, p, 0, e)
print("wavefront", ctx.stats())
