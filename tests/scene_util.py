"""Synthetic scenes shared by the tests and bench.py (test/bench infrastructure, not product code)."""
import math
import os

import numpy as np


def write_heightfield_obj(path, n=225, amp=0.1):
    """n x n vertex grid over [-1,1]^2, z = amp*sin(6 pi u)*cos(6 pi v), 2*(n-1)^2 triangles,
    one `g`, written as v/f text so that it goes through the OBJ import path (and receives main.rs:802's transform)."""
    u = np.linspace(-1.0, 1.0, n)
    with open(path, "w") as f:
        f.write("# synthetic height field\ng Heightfield\n")
        for j in range(n):
            for i in range(n):
                z = amp * math.sin(6 * math.pi * u[i]) * math.cos(6 * math.pi * u[j])
                f.write(f"v {u[i]:.6f} {u[j]:.6f} {z:.6f}\n")
        for j in range(n - 1):
            for i in range(n - 1):
                a = j * n + i + 1
                b, c, d = a + 1, a + n, a + n + 1
                f.write(f"f {a} {b} {d}\nf {a} {d} {c}\n")
    return 2 * (n - 1) * (n - 1)


def fixture_plus_mesh(b, tmp_dir, n=225):
    """The scene literal of main() plus the synthetic mesh as extra triangles of object 0 (C5)."""
    path = os.path.join(str(tmp_dir), f"heightfield_{n}.obj")
    ntri = write_heightfield_obj(path, n)
    w = b.World.fixture()
    got = b.ObjectProxy(w, 0).load_obj(path)          # same /3 + (0.7, 1.0, -0.5) transform as load_obj
    assert got == ntri
    return w, ntri
