import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_binding as ob
b = ob.b
ctx = b.Context(0); w = b.World.fixture(); ctx.upload_scene(w); cam = b.fixture_camera()
p = b.default_params(width=3840, height=2160)
rgb, prim = ctx.render_whitted(cam, p)
bad = ~np.isfinite(rgb).all(axis=2)
ys, xs = np.where(bad)
print("non-finite pixels:", len(ys), list(zip(ys[:10], xs[:10])), "prims", prim[ys[:10], xs[:10]])
for y in sorted(set(ys.tolist()))[:4]:
    band = b.copy_params(p, row_begin=int(y), row_count=1)
    o_rgb, o_prim, _ = ob.render_whitted(w.scene(), cam, band)
    xs_y = xs[ys == y]
    print("row", y, "gpu", rgb[y, xs_y[:3]], "oracle", o_rgb[y, xs_y[:3]], "oracle nonfinite count in row", (~np.isfinite(o_rgb[y]).all(axis=1)).sum(), "same positions", np.array_equal(~np.isfinite(o_rgb[y]).all(axis=1), bad[y]))
