"""Dev: time the wavefront tracer (CUDA events inside the library) for the library named by B200RT_LIB."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
b = g.load_package()
ctx = b.Context(0)
ctx.upload_scene(b.World.fixture())
cam = b.fixture_camera()
w, h, e = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "3840x2160x4").split("x"))
tr = b.TRACER_MEGAKERNEL if (len(sys.argv) > 2 and sys.argv[2] == "mega") else b.TRACER_WAVEFRONT
p = b.default_params(width=w, height=h, seed=0, tracer=tr)
ctx.render_distributed(cam, p, 0, e)
ms = []
for _ in range(3):
    ctx.render_distributed(cam, p, 0, e)
    ms.append(ctx.stats()["kernel_ms"])
print(os.environ.get("B200RT_LIB", "default"), f"{w}x{h}x{e}", "kernel_ms", [round(m, 2) for m in ms], "Mrays/s", round(w * h * e / min(ms) / 1e3, 1), flush=True)
