"""The C-ABI library: loads, exports every symbol include/b200rt.h declares, PODs have the documented layout,
and WITHOUT a GPU every compute entry point fails loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200rt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(b200rt):
    lib = b200rt.load_library()
    syms = declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(b200rt.EXPORTED_SYMBOLS) == syms       # the binding's list is the header's list


def test_pod_layouts(b200rt):
    # geometric.rs:43-47 is 8 floats; Triangle = 3 vertices + index = 100 bytes
    assert C.sizeof(b200rt.Vertex) == 32
    assert C.sizeof(b200rt.Triangle) == 100
    assert C.sizeof(b200rt.Sphere) == 20
    assert C.sizeof(b200rt.Material) == 4 + 12 + 12 + 4 + 12 + 16 + 8 + 32
    assert C.sizeof(b200rt.Light) == 8 + 12 + 12 + 8 + 12
    assert C.sizeof(b200rt.Camera) == 44
    assert C.sizeof(b200rt.Ray) == 36 and C.sizeof(b200rt.Hit) == 48
    assert C.sizeof(b200rt.Params) == 56
    assert b200rt.Params.seed.offset == 40


def test_defaults_are_the_reference_literals(b200rt):
    p = b200rt.default_params()
    assert (p.width, p.height, p.depth, p.tir_retries) == (1280, 960, 5, 10)        # main.rs:1084-1085, 1098, 378
    assert p.threshold == pytest.approx(0.001) and p.refract_max_distance == 100.0     # main.rs:467, 505
    assert p.focus == 3.0 and p.blur == pytest.approx(0.04)                            # main.rs:1147-1148
    cam = b200rt.fixture_camera()
    np.testing.assert_allclose(list(cam.center), [2.0, 2.5, 2.0])
    np.testing.assert_allclose(list(cam.toward), [-3 ** -0.5] * 3, rtol=1e-6)
    assert cam.near == pytest.approx(-0.1) and cam.fovy == pytest.approx(np.pi / 3, rel=1e-6)


def test_strerror(b200rt):
    assert b200rt.strerror(0) == "ok"
    assert "CPU fallback" in b200rt.strerror(b200rt.ERR_NO_DEVICE)
    assert b200rt.strerror(-999) == "unknown error"


def test_null_arguments_are_rejected(b200rt):
    lib = b200rt.load_library()
    assert lib.b200rt_create(0, None) == b200rt.ERR_INVALID
    assert lib.b200rt_destroy(None) == b200rt.ERR_INVALID
    assert lib.b200rt_upload_scene(None, None) == b200rt.ERR_INVALID
    assert lib.b200rt_world_push_object(None, None) == b200rt.ERR_INVALID
    assert lib.b200rt_render_whitted(None, None, None, None, None) == b200rt.ERR_INVALID


def test_no_gpu_means_no_result(b200rt):
    """On a machine without a CUDA device the product path must fail, not fall back to a CPU renderer."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(b200rt.B200rtError) as e:
        b200rt.Context(0)
    assert e.value.code == b200rt.ERR_NO_DEVICE


def test_product_does_not_reference_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may touch oracle/."""
    pkg = os.path.join(ROOT, "homework-18-graphics-raytracer_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_lib" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_binding" not in text and "liboracle" not in text and "oracle.h" not in text, f
