"""Writes tests/golden/fixture_scene.npz: the reference's scene literal (main.rs:810-1075), camera (main.rs:1077-1083)
and default render parameters as the raw C-ABI records `World.fixture()` produces.

    python tests/golden/make_fixture_scene.py

bench.py's `--impl reference` arm (the CPU oracle on the host cores) builds its scene from this file, so that arm
never loads libb200rt.so; tests/test_host_world.py checks that the file equals `World.fixture()` byte for byte.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import __graft_entry__ as ge  # noqa: E402


def raw(ptr, n, typ):
    return np.frombuffer(C.string_at(ptr, n * C.sizeof(typ)), dtype=np.uint8).copy()


def main():
    b = ge.load_package()
    w = b.World.fixture()
    s = w.scene()
    cam = b.fixture_camera()
    p = b.default_params()
    np.savez(os.path.join(HERE, "fixture_scene.npz"),
             triangles=raw(s.triangles, s.n_triangles, b.Triangle), spheres=raw(s.spheres, s.n_spheres, b.Sphere),
             materials=raw(s.materials, s.n_materials, b.Material), lights=raw(s.lights, s.n_lights, b.Light),
             camera=np.frombuffer(bytes(cam), dtype=np.uint8).copy(), params=np.frombuffer(bytes(p), dtype=np.uint8).copy())
    print("wrote fixture_scene.npz:", s.n_triangles, "triangles,", s.n_spheres, "spheres,", s.n_materials, "materials,", s.n_lights, "lights")


if __name__ == "__main__":
    main()
