"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the metric accumulation of the reference's training and
evaluation loops.

Follows main.py:110-143 (train) / :175-199 (evaluate): y and y_pred are de-normalised with
NPZSequenceDataset.denormalize (unet.py:306-327), `diff = pred - y`, the valid pixels are `mask.astype(bool)`
(all pixels without a mask), and MAE = mean |diff|, RMSE = sqrt(mean diff^2), ME = mean diff over ALL valid pixels
of the epoch (zeros when there is none).  Pinned by tests/golden/metrics_main_evaluate.npz, produced by the
reference's own `evaluate` (tests/golden/make_golden_metrics.py).  Only tests/ may import this module.
"""
import numpy as np


def denormalize(y_norm, trans_min, trans_max, y_scale, y_transform):
    yt = (np.asarray(y_norm, dtype=np.float64) + 1.0) / 2.0 * (trans_max - trans_min) + trans_min   # unet.py:316
    if y_transform == "asinh":
        return np.sinh(yt) * y_scale                                                                 # unet.py:319
    if y_transform == "signed_log":
        return np.sign(yt) * (np.expm1(np.abs(yt)) * y_scale)                                        # unet.py:321
    return yt


def epoch_metrics(batches, trans_min, trans_max, y_scale, y_transform, use_mask=True):
    """batches: iterable of (y_pred, y, mask-or-None).  Returns (mae, rmse, me)."""
    s_abs = s_sq = s_err = 0.0
    cnt = 0
    for yp, y, mask in batches:
        d = (denormalize(yp, trans_min, trans_max, y_scale, y_transform)
             - denormalize(y, trans_min, trans_max, y_scale, y_transform))
        if use_mask and mask is not None:
            d = d[np.asarray(mask).astype(bool)]                                                     # main.py:122-125
        d = d.ravel()
        s_abs += np.abs(d).sum()
        s_sq += (d ** 2).sum()
        s_err += d.sum()
        cnt += d.size
    if cnt == 0:
        return 0.0, 0.0, 0.0                                                                         # main.py:142-143
    return s_abs / cnt, float(np.sqrt(s_sq / cnt)), s_err / cnt
