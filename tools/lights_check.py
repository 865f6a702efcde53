import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import __graft_entry__ as g
b = g.load_package()
import oracle_binding as ob
cam = b.fixture_camera()
w = b.World.fixture()
for k in range(3):
    w.push_light(b.point_light([1.0 + k, 2.0, 1.5 - k], [0.3, 0.4, 0.5]))
ctx = b.Context(0)
ctx.upload_scene(w)
p = b.default_params(width=160, height=120, seed=4)
acc = ctx.render_distributed(cam, p, 0, 2)
mega = ctx.render_distributed(cam, b.copy_params(p, tracer=b.TRACER_MEGAKERNEL), 0, 2)
o_acc, _ = ob.render_distributed(w.scene(), cam, p, 0, 2)
print("wave==mega counts", np.array_equal(acc[..., 3], mega[..., 3]), "n diff", int((acc[..., 3] != mega[..., 3]).sum()))
print("wave==oracle counts", np.array_equal(acc[..., 3], o_acc[..., 3]), "n diff", int((acc[..., 3] != o_acc[..., 3]).sum()))
print("mega==oracle counts", np.array_equal(mega[..., 3], o_acc[..., 3]), "n diff", int((mega[..., 3] != o_acc[..., 3]).sum()))
print("max abs wave-mega", float(np.nanmax(np.abs(acc - mega))))
d = np.argwhere(acc[..., 3] != o_acc[..., 3])[:5]
for y, x in d: print(y, x, acc[y, x], mega[y, x], o_acc[y, x])
