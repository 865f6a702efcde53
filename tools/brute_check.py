"""Dev check: the two-phase cast against the brute-force exact cast inside the stochastic tracer (bitwise)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
b = g.load_package()
ctx = b.Context(0)
ctx.upload_scene(b.World.fixture())
cam = b.fixture_camera()
w, h, ep = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "3840x2160x4").split("x"))
res = {}
for name, cm, tr in (("two_phase/mega", b.CAST_TWO_PHASE, b.TRACER_MEGAKERNEL), ("brute/mega", b.CAST_BRUTE_EXACT, b.TRACER_MEGAKERNEL),
                     ("two_phase/wave", b.CAST_TWO_PHASE, b.TRACER_WAVEFRONT)):
    p = b.default_params(width=w, height=h, seed=0, tracer=tr, cast_mode=cm)
    ctx.reset_stats()
    acc = ctx.render_distributed(cam, p, 0, ep)
    st = ctx.stats()
    res[name] = acc
    print(name, "casts", st["casts"], "samples", st["samples"], "kernel_ms", st["kernel_ms"], flush=True)
ref = res["brute/mega"]
for name in ("two_phase/mega", "two_phase/wave"):
    a = res[name]
    diff = (a.view(np.uint32) != ref.view(np.uint32)).any(axis=2)
    print(name, "vs brute: bitwise equal", not diff.any(), "pixels differing", int(diff.sum()), np.argwhere(diff)[:8].tolist())
