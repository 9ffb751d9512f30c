"""Encoding-score scoring on the device.

The ridge fit of `visreps/analysis/encoding_score.py` stays with himalaya; what runs here is what follows the
fit: `correlation_score(Y_test, pred)` (himalaya.scoring, per-voxel Pearson r, encoding_score.py:61, :214) and
the 90 %-subsample bootstrap over the cached predictions (encoding_score.py:228-240), which draws from the same
`RandomState` the function created for its train split (encoding_score.py:107, :131, :233).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib, rsa


def _as_device_f32(t, dev: torch.device) -> torch.Tensor:
    t = torch.as_tensor(t)
    if t.ndim != 2:
        raise ValueError("y_true and y_pred must be 2-D (n_samples, n_targets)")
    t = t.to(device=dev, dtype=torch.float32)
    return t if t.stride(1) == 1 else t.contiguous()


def _column_scores(y_true, y_pred, masks: torch.Tensor, per_column: bool):
    lib = _lib.load()
    dev = masks.device
    yt, yp = _as_device_f32(y_true, dev), _as_device_f32(y_pred, dev)
    if yt.shape != yp.shape:
        raise ValueError("y_true and y_pred must have the same shape")
    n, V = int(yt.shape[0]), int(yt.shape[1])
    R = int(masks.shape[0])
    mean = torch.empty(R, dtype=torch.float64, device=dev)
    cols = torch.empty((R, V), dtype=torch.float32, device=dev) if per_column else None
    ws = _lib.workspace(lib.vr_bootstrap_column_corr_workspace_bytes(n, V, R), dev)
    _lib.check(lib.vr_bootstrap_column_corr_f32(yt.data_ptr(), yp.data_ptr(), n, V, yt.stride(0), yp.stride(0), masks.data_ptr(), R,
                                                mean.data_ptr(), cols.data_ptr() if cols is not None else None,
                                                ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "vr_bootstrap_column_corr_f32")
    return mean, cols


def correlation_score(y_true, y_pred) -> torch.Tensor:
    """himalaya.scoring.correlation_score: Pearson r of every column pair, shape (n_targets,), on the device."""
    dev = _lib.require_cuda()
    n = int(torch.as_tensor(y_true).shape[0])
    _, cols = _column_scores(y_true, y_pred, rsa.build_masks(None, n, dev, with_full=True), per_column=True)
    return cols[0]


def bootstrap_encoding_score(y_true, y_pred, n_bootstrap: int = 1000, seed: int = 42,
                             rng: Optional[np.random.RandomState] = None) -> Tuple[float, float, float, List[float]]:
    """Point estimate (mean column correlation over all samples, encoding_score.py:61) and the bootstrap of
    encoding_score.py:228-240: `n_bootstrap` draws of int(0.9 n) samples without replacement from `rng`
    (default RandomState(seed)), each scored by the mean column correlation; 2.5 / 97.5 percentiles.
    Returns (point_estimate, ci_low, ci_high, bootstrap_scores) with the scores in draw order."""
    dev = _lib.require_cuda()
    n = int(torch.as_tensor(y_true).shape[0])
    draws = rsa.draw_subsample_masks(n, n_bootstrap, seed, rng)
    masks = rsa.build_masks(draws, n, dev, with_full=True)
    mean, _ = _column_scores(y_true, y_pred, masks, per_column=False)
    out = mean.cpu().numpy()
    scores = out[1:]
    lo, hi = rsa.percentile_ci(scores)
    return float(out[0]), float(lo), float(hi), scores.tolist()
