"""write_to_file (main.rs:764-776): the host-side PNG writer of the C ABI — decoded pixels round-trip,
the image appears under its final name only (tmp.png + rename), errors are reported as codes."""
import os

import numpy as np
import pytest


@pytest.mark.parametrize("size", [(1, 1), (7, 5), (640, 480), (21846, 1), (300, 233)])
def test_png_round_trip(b200rt, tmp_path, size):
    from PIL import Image
    w, h = size
    img = np.random.default_rng(w * 31 + h).integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    out = tmp_path / "out.png"
    b200rt.write_png(str(out), img)
    back = np.asarray(Image.open(out).convert("RGB"))
    assert back.shape == img.shape and np.array_equal(back, img)
    assert not (tmp_path / "tmp.png").exists()              # renamed over the target, main.rs:774-775
    # overwriting an existing image is the per-epoch path of the reference (main.rs:1172)
    b200rt.write_png(str(out), 255 - img)
    assert np.array_equal(np.asarray(Image.open(out).convert("RGB")), 255 - img)


def test_png_header_is_rgb8(b200rt, tmp_path):
    out = tmp_path / "a.png"
    b200rt.write_png(str(out), np.zeros((3, 4, 3), dtype=np.uint8))
    raw = out.read_bytes()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n" and raw[12:16] == b"IHDR"
    assert int.from_bytes(raw[16:20], "big") == 4 and int.from_bytes(raw[20:24], "big") == 3
    assert raw[24] == 8 and raw[25] == 2                    # 8 bits, ColorType::RGB (main.rs:770)


def test_png_errors(b200rt, tmp_path):
    with pytest.raises(b200rt.B200rtError):
        b200rt.write_png(str(tmp_path / "no_such_dir" / "x.png"), np.zeros((2, 2, 3), dtype=np.uint8))
    with pytest.raises(ValueError):
        b200rt.write_png(str(tmp_path / "x.png"), np.zeros((2, 2), dtype=np.uint8))
    lib = b200rt.load_library()
    assert lib.b200rt_write_png_rgb8(None, None, 0, 0) < 0
