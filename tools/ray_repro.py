"""Dev: cast specific rays (bit patterns) through two-phase and brute-exact b200rt_intersect and the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import __graft_entry__ as g
b = g.load_package()
import oracle_binding as ob
ctx = b.Context(0)
world = b.World.fixture()
ctx.upload_scene(world)
cases = [
    ("3f959753 00000000 bf87819a", "bf00d52e 3f3c2e41 bee89aa4", 0, 37, 1),
    ("40042e8a 4024e36a 400203d6", "befd3bb2 bead9492 bf4cdeb7", 0, -1, 0),
    ("3f83e17c 00000000 bf0a1ae0", "beaf1a07 3f4ca2b9 befcf18c", 0, 37, 1),
    ("bfd6e906 00000000 3fb82ed9", "3f7b6335 3e249b5d bdcb772a", 0, 36, 1),
]
rays = np.zeros(len(cases) * 32, dtype=b.RAY_DTYPE)
for i in range(len(rays)):
    o, d, face, ex, exf = cases[i % len(cases)]
    rays["origin"][i] = np.array([int(x, 16) for x in o.split()], dtype=np.uint32).view(np.float32)
    rays["direction"][i] = np.array([int(x, 16) for x in d.split()], dtype=np.uint32).view(np.float32)
    rays["face_direction"][i] = face; rays["exclude_prim"][i] = ex; rays["exclude_face"][i] = exf
t = ctx.intersect(rays, b.CAST_TWO_PHASE)
e = ctx.intersect(rays, b.CAST_BRUTE_EXACT)
o = ob.intersect(world.scene(), rays)
for i in range(len(cases)):
    print(i, "two_phase", t["prim_id"][i], t["distance"][i], "| brute", e["prim_id"][i], e["distance"][i], "| oracle", o["prim_id"][i], o["distance"][i])
print(ctx.stats())
