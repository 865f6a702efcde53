"""Pins the CPU oracle to the reference's own artefacts.

The reference has no numeric golden vectors; its only checked-in outputs are PNGs.  The deterministic
pass's render (report/out_single_epoch.png) is the strongest pin there is: the oracle must reproduce it
after the same post_process + sRGB/u8 encoding.  Fixture: tests/golden/out_single_epoch_probe.json
(made by tests/golden/make_golden.py from the reference tree)."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def oracle_frame(b200rt, oracle, fixture_world):
    cam = b200rt.fixture_camera()
    params = b200rt.default_params()          # 1280x960, depth 5: the reference's own frame
    rgb, prim, cnt = oracle.render_whitted(fixture_world.scene(), cam, params)
    return rgb, prim, cnt


def test_oracle_reproduces_reference_png(oracle, oracle_frame):
    gold = json.load(open(os.path.join(GOLD, "out_single_epoch_probe.json")))
    rgb, prim, _ = oracle_frame
    assert rgb.shape == (gold["height"], gold["width"], 3)
    pp, p99 = oracle.post_process(rgb)
    img = oracle.encode_srgb8(pp).astype(int)
    probes = np.array(gold["probes_y_x_r_g_b"])
    got = img[probes[:, 0], probes[:, 1]]
    diff = np.abs(got - probes[:, 2:5]).max(axis=1)
    # 8-bit sRGB after a data-dependent global scale: +-2 levels; isolated silhouette pixels may differ
    # (7 of 1 228 800 over the full image when compared in the authoring container)
    assert (diff <= 2).mean() >= 0.998, f"{(diff > 2).sum()} of {len(diff)} probes off by more than 2/255"
    assert np.median(diff) <= 1
    # global statistics of the frame
    # near-black pixels quantise to 0 or 1 depending on the global p99 scale (ours 0.834, the image's ~0.841)
    assert abs(int((img.max(axis=2) == 0).sum()) - gold["black_pixels"]) <= 0.02 * gold["black_pixels"]
    assert abs(int((img.max(axis=2) == 255).sum()) - gold["saturated_pixels"]) <= 0.05 * gold["saturated_pixels"]
    np.testing.assert_allclose(img.reshape(-1, 3).mean(axis=0), gold["mean_rgb"], atol=1.0)


def test_oracle_full_image_against_reference_tree(oracle, oracle_frame):
    """Whole-image comparison, only where the reference tree is mounted (authoring container)."""
    path = "/root/reference/report/out_single_epoch.png"
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted")
    from PIL import Image
    ref = np.array(Image.open(path).convert("RGB")).astype(int)
    rgb, _, _ = oracle_frame
    pp, _ = oracle.post_process(rgb)
    img = oracle.encode_srgb8(pp).astype(int)
    d = np.abs(img - ref).max(axis=2)
    assert (d > 2).sum() <= 16, (d > 2).sum()
    assert d.mean() < 1.0


def test_primary_hit_id_census(oracle_frame, fixture_world):
    """Every primitive class of the scene literal is visible; ids follow push order (main.rs:183, 264)."""
    _, prim, cnt = oracle_frame
    ids = set(np.unique(prim).tolist())
    assert -1 in ids                      # background
    assert ids & set(range(0, 36))        # dodecahedron (OBJ import)
    assert {36, 37} <= ids                # floor
    assert {38, 39} & ids                 # textured wall
    assert ids & set(range(40, 64))       # glass slabs
    assert {64, 65, 66, 67} <= ids        # four spheres
    assert cnt["samples"] == 1280 * 960
    assert cnt["tri_pairs"] == cnt["casts"] * 64 and cnt["sph_pairs"] == cnt["casts"] * 4


def test_dodecahedron_obj_roundtrip(b200rt, tmp_path):
    """load_obj (main.rs:778-807): the built-in mesh equals dodecahedron.obj loaded through the importer."""
    mesh = json.load(open(os.path.join(GOLD, "dodeca_mesh.json")))
    obj = tmp_path / "dodecahedron.obj"
    with open(obj, "w") as f:
        f.write("# written by the test from tests/golden/dodeca_mesh.json\ng Object001\n\n")
        for v in mesh["v"]:
            f.write("v  " + "  ".join(v) + "\n")
        f.write("\n")
        for face in mesh["f"]:
            f.write("f  " + "  ".join(str(i) for i in face) + "\n")
    a = b200rt.World.fixture()
    bw = b200rt.World.fixture(str(obj))
    sa, sb = a.scene(), bw.scene()
    assert sa.n_triangles == sb.n_triangles == 64
    import ctypes as C
    raw_a = C.string_at(sa.triangles, C.sizeof(b200rt.Triangle) * 64)
    raw_b = C.string_at(sb.triangles, C.sizeof(b200rt.Triangle) * 64)
    assert raw_a == raw_b
    if os.path.exists("/root/reference/dodecahedron.obj"):
        c = b200rt.World.fixture("/root/reference/dodecahedron.obj")
        raw_c = C.string_at(c.scene().triangles, C.sizeof(b200rt.Triangle) * 64)
        assert raw_a == raw_c
    # every source vertex has |v| ~ 1, so after p/3 + (0.7,1.0,-0.5) (main.rs:802) it is 1/3 from that centre
    tri = np.frombuffer(raw_a, dtype=np.float32).reshape(64, 25)[:36]
    pos = tri[:, [0, 1, 2, 8, 9, 10, 16, 17, 18]].reshape(-1, 3)
    dist = np.linalg.norm(pos - np.array([0.7, 1.0, -0.5], dtype=np.float32), axis=1)
    np.testing.assert_allclose(dist, 1.0 / 3.0, atol=2e-6)
