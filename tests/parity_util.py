"""Agreement of a stochastic render {sum.rgb, count} with the oracle's, shared by the GPU tests and smoke().

Both sides draw the same Philox sample streams, so every pixel sample is the same path unless a libm difference of a few
ulps (CUDA vs glibc sinf / cosf / acosf / powf / logf in scatter_hit and the lens) moves a scattered ray across a
silhouette.  Three numbers describe that:
  frac_moved   fraction of pixels whose accumulated colour differs by more than 1e-3 relative (a moved sample)
  max_abs      the largest absolute difference of an accumulated channel (one moved sample changes a pixel by at most one
               sample's radiance, so it is bounded by a few times the frame's peak)
  psnr         PSNR of the mean images sum / count against the oracle's p99 luma (the north_star's stochastic criterion)
"""
import numpy as np

ABS_FLOOR = 1e-3
LUMA = np.array([0.2126729, 0.7151522, 0.0721750], dtype=np.float32)


def stochastic_agreement(acc, o_acc):
    acc = np.asarray(acc, dtype=np.float32)
    o_acc = np.asarray(o_acc, dtype=np.float32)
    count_diff = int((acc[..., 3] != o_acc[..., 3]).sum())
    a, o = acc[..., :3], o_acc[..., :3]
    fin = np.isfinite(a).all(axis=-1) & np.isfinite(o).all(axis=-1)
    err = np.abs(a - o) / np.maximum(np.maximum(np.abs(a), np.abs(o)), ABS_FLOOR)
    err = np.where(fin[..., None], err, 0.0)
    mean_g = np.where(acc[..., 3:4] > 0, a / np.maximum(acc[..., 3:4], 1), 0.0)
    mean_o = np.where(o_acc[..., 3:4] > 0, o / np.maximum(o_acc[..., 3:4], 1), 0.0)
    mean_g = np.where(fin[..., None], mean_g, 0.0)
    mean_o = np.where(fin[..., None], mean_o, 0.0)
    peak = float(np.percentile(mean_o @ LUMA, 99))
    mse = float(np.mean((mean_g - mean_o) ** 2))
    return {
        "count_diff": count_diff,
        "frac_moved": float((err.max(axis=-1) > 1e-3).mean()),
        "max_abs": float(np.where(fin[..., None], np.abs(a - o), 0.0).max()) if a.size else 0.0,
        "peak": peak,
        "psnr": float(10 * np.log10(peak * peak / max(mse, 1e-30))) if peak > 0 else float("inf"),
        "nonfinite_mismatch": int((np.isfinite(a).all(axis=-1) != np.isfinite(o).all(axis=-1)).sum()),
    }
