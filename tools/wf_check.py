"""Dev check: wavefront vs megakernel stochastic tracer — bitwise equality of the accumulators and timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
b = g.load_package()
ctx = b.Context(0)
ctx.upload_scene(b.World.fixture())
cam = b.fixture_camera()
sizes = [(64, 48, 3), (320, 240, 4), (960, 540, 4)]
if len(sys.argv) > 1:
    sizes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (w, h, ep) in sizes:
    res = {}
    for name, tr in (("mega", b.TRACER_MEGAKERNEL), ("wave", b.TRACER_WAVEFRONT)):
        p = b.default_params(width=w, height=h, seed=7, tracer=tr)
        ctx.render_distributed(cam, p, 0, ep)      # warm-up (module load, workspace allocation)
        ctx.reset_stats()
        t0 = time.perf_counter()
        acc = ctx.render_distributed(cam, p, 0, ep)
        dt = time.perf_counter() - t0
        st = ctx.stats()
        res[name] = (acc, st, dt)
        print(f"{w}x{h}x{ep} {name}: kernel_ms {st['kernel_ms']:.3f} wall {dt*1e3:.1f} ms casts {st['casts']} samples {st['samples']} confirms {st['exact_confirms']} fallbacks {st['certify_fallbacks']} rounds {st['wavefront_rounds']}", flush=True)
    a, m = res["wave"][0], res["mega"][0]
    same = np.array_equal(a.view(np.uint32), m.view(np.uint32))
    print("  bitwise equal:", same, " count equal:", np.array_equal(a[..., 3], m[..., 3]),
          " max abs diff:", float(np.nanmax(np.abs(a - m))), " casts equal:", res["wave"][1]["casts"] == res["mega"][1]["casts"])
