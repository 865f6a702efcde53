"""Generates tests/golden/stochastic_pin.npz from the reference's own renders of the STOCHASTIC pass
(run in the authoring container, where /root/reference exists):

    python tests/golden/make_stochastic_pin.py

report/out.png is what main() leaves after its 100 epochs (main.rs:1129-1173: thin-lens depth of field, focus 3.0,
blur 0.04, scatter tracer, depth 5, 1280x960); report/out_small_blur.png is the same loop with a smaller lens (its
blur literal is not recorded in the reference tree).  Every epoch ends with post_process (main.rs:1171), so the file is
img_k = (img_{k-1} + samples_k) / p99_k: a geometrically weighted sum of the last few epochs' samples - a global,
random SCALE times a noisy estimate of the mean sample image (about three effective samples per pixel).  The fixture
keeps what a test can compare with the oracle's mean image, up to that one scale:
  * blocks32 / valid32   linear-light means of the 40 x 30 blocks of 32 x 32 pixels (valid = under 2 % of the block's
                         values clipped at 255 by the u8 encode, image.rs:63);
  * sens_idx / sens4     4 x 4-pixel block means (= the pixels of a 320 x 240 frame of the same camera: clip_x, clip_y of
                         main.rs:1094-1095 are the same for pixel 4x, 4y) at the 10 % of the pixels whose oracle mean
                         changes most between blur 0 and blur 0.12 - the pixels that see the lens.
The oracle renders that choose those pixels are part of this script (seed 0, 32 epochs), so the fixture is reproducible.
"""
import os
import sys

import numpy as np
from PIL import Image

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_binding as ob  # noqa: E402


def srgb_to_linear(u8):
    c = u8.astype(np.float64) / 255.0
    return np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)


def blocks(img, b):
    h, w, c = img.shape
    return img.reshape(h // b, b, w // b, b, c).mean(axis=(1, 3))


def main():
    fx = ob.GoldenFixture()
    W, H, E = 320, 240, 32

    def mean_img(blur):
        p = fx.params(width=W, height=H, depth=5, seed=0, focus=3.0, blur=blur)
        acc, _ = ob.render_distributed(fx.scene, fx.camera, p, 0, E, n_threads=ob.host_threads())
        return acc[..., :3] / E

    sens = np.abs(mean_img(0.0) - mean_img(0.12)).sum(axis=2)
    out = {}
    for key, name in (("out", "out.png"), ("small", "out_small_blur.png")):
        raw = np.array(Image.open(os.path.join(REF, "report", name)).convert("RGB"))
        assert raw.shape == (960, 1280, 3)
        lin = srgb_to_linear(raw)
        clipped = (raw >= 255).astype(np.float64)
        out[f"{key}_blocks32"] = blocks(lin, 32).astype(np.float32)
        out[f"{key}_valid32"] = blocks(clipped, 32).max(axis=2) < 0.02
        unclipped4 = blocks(clipped, 4).max(axis=2) == 0
        mask = (sens > np.quantile(sens, 0.90)) & unclipped4
        idx = np.flatnonzero(mask.reshape(-1)).astype(np.uint32)
        out[f"{key}_sens_idx"] = idx
        out[f"{key}_sens4"] = blocks(lin, 4).reshape(-1, 3)[idx].astype(np.float16)
    np.savez_compressed(os.path.join(HERE, "stochastic_pin.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    main()
