"""The oracle's STOCHASTIC pass (shoot_focus main.rs:101-127, distributed_ray_trace 521-614, scatter_hit 539-554,
weighted_select 652-666, the is_normal filter 1157-1160) pinned to the reference's own renders of that pass:
report/out.png (what main() leaves after its 100 epochs: focus 3.0, blur 0.04) and report/out_small_blur.png.

What the files are: every epoch ends with post_process (main.rs:1171), img_k = (img_{k-1} + samples_k) / p99_k, so the
file is a geometrically weighted sum of the last few epochs' samples = ONE global random scale x a noisy estimate
(about three effective samples per pixel) of the mean sample image, then sRGB-u8 encoded (values above 1 clip).  The
random-number generator of the reference (rand 0.5 ISAAC + ziggurat) is not reproduced, so the comparison is the one
the north_star prescribes for stochastic scenes - converged means - up to that scale:

  * 32 x 32-pixel block means (1024 pixels x ~3 samples: reference noise ~1 %) against the oracle's mean image of 32
    epochs at 320 x 240 (same camera rays at every fourth pixel; an 8 x 8 block holds 2048 samples: noise ~1 %):
    measured relative RMS residual 4.0 % (bound 6 %), correlation 0.9975 (bound 0.995), and the three channels' scales
    agree to 0.1 % (bound 1 %): the energy split between the shading / scatter branches and the colour balance.
  * the lens: at the tenth of the pixels whose mean changes most between blur 0 and blur 0.12, the residual against
    out.png is smallest for the blur the code states (0.04) - measured 0.33 against 0.37 (no lens) and 0.43 (0.12).

The fixture (tests/golden/stochastic_pin.npz) is generated from /root/reference by tests/golden/make_stochastic_pin.py.
"""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 320, 240


@pytest.fixture(scope="module")
def pin():
    return np.load(os.path.join(ROOT, "tests", "golden", "stochastic_pin.npz"))


def oracle_mean(oracle, blur, epochs):
    fx = oracle.GoldenFixture()
    p = fx.params(width=W, height=H, depth=5, seed=0, focus=3.0, blur=blur)
    acc, cnt = oracle.render_distributed(fx.scene, fx.camera, p, 0, epochs, n_threads=oracle.host_threads())
    # the reference adds the accepted samples of an epoch and nothing for the dropped ones (main.rs:1157-1167):
    # the expectation of one epoch's contribution is sum / epochs, not sum / accepted
    return acc[..., :3].astype(np.float64) / epochs, cnt


def blocks(img, b):
    h, w, c = img.shape
    return img.reshape(h // b, b, w // b, b, c).mean(axis=(1, 3))


def fit(x, y):
    """least-squares scale a of y ~ a x, relative RMS residual, correlation"""
    a = float((x * y).sum() / (x * x).sum())
    return a, float(np.sqrt(((y - a * x) ** 2).sum() / (y ** 2).sum())), float(np.corrcoef(x.ravel(), y.ravel())[0, 1])


@pytest.fixture(scope="module")
def mean_004(oracle):
    return oracle_mean(oracle, 0.04, 32)[0]


@pytest.mark.parametrize("key", ["out", "small"])
def test_block_means_match_the_reference_render(pin, mean_004, key):
    ref, valid = pin[f"{key}_blocks32"].astype(np.float64), pin[f"{key}_valid32"]
    assert valid.sum() > 1000                                   # of 1200 blocks (the others hold clipped highlights)
    mine = blocks(mean_004, 8)
    a, resid, corr = fit(mine[valid], ref[valid])
    scales = [fit(mine[valid][:, c], ref[valid][:, c])[0] for c in range(3)]
    print(f"{key}: scale {a:.4f} (r/g/b {scales[0]:.4f} {scales[1]:.4f} {scales[2]:.4f}) residual {resid:.4f} corr {corr:.5f}")
    assert resid <= 0.06, resid
    assert corr >= 0.995, corr
    assert max(scales) / min(scales) <= 1.01, scales           # colour balance of the whole pass
    assert 0.5 < a < 1.5                                        # the chain's normaliser stays near the p99 of one epoch


def test_the_lens_blur_of_out_png_is_the_code_literal(pin, oracle):
    """shoot_focus: main.rs:1147-1148 passes focus 3.0 and blur 0.04."""
    idx, ref = pin["out_sens_idx"], pin["out_sens4"].astype(np.float64)
    resid = {}
    for blur in (0.0, 0.04, 0.12):
        mine = oracle_mean(oracle, blur, 16)[0].reshape(-1, 3)[idx]
        resid[blur] = fit(mine, ref)[1]
    print("residual at the lens-sensitive pixels:", resid)
    assert resid[0.04] < resid[0.0] - 0.015 and resid[0.04] < resid[0.12] - 0.05, resid
