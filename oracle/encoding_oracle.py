"""CPU oracle for the encoding-score bootstrap  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Restates  /root/reference/visreps/analysis/encoding_score.py:228-240  (the 90 %-subsample bootstrap over cached
test predictions) and the one third-party function it calls, `himalaya.scoring.correlation_score`.

PARITY UNPINNED for the third-party part: himalaya (pinned 0.4.9 in the reference's uv.lock:965-966) is not
installed in this image and there is no network, so `correlation_score` is restated from its published
definition -- the Pearson correlation of every column pair, computed as the mean over samples of the product of
population z-scores -- and cross-checked against `scipy.stats.pearsonr` per column in
tests/test_oracle_golden.py.  The bootstrap loop itself (RNG stream, subsample size, percentiles) is the same
numpy code the RSA oracle is pinned on.  The reference has no test for this block.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def correlation_score(y_true: np.ndarray, y_pred: np.ndarray) -> np.ndarray:
    """himalaya.scoring.correlation_score(y_true, y_pred): (n_samples, n_targets) x2 -> (n_targets,).
    Evaluated in float32 like the reference (torch_cuda backend, float32 tensors)."""
    yt = np.asarray(y_true, dtype=np.float32)
    yp = np.asarray(y_pred, dtype=np.float32)

    def zscore(x):
        x = x - x.mean(0, keepdims=True)
        return x / np.sqrt((x ** 2).mean(0, keepdims=True))

    with np.errstate(invalid="ignore", divide="ignore"):
        return (zscore(yt) * zscore(yp)).mean(0)


def bootstrap_encoding(y_true: np.ndarray, y_pred: np.ndarray, n_bootstrap: int = 1000, seed: int = 42,
                       rng: Optional[np.random.RandomState] = None):
    """encoding_score.py:61 (point estimate) and :228-240 (bootstrap) -> (point, ci_low, ci_high, scores)."""
    rng = np.random.RandomState(seed) if rng is None else rng
    n_test = y_true.shape[0]
    point = float(correlation_score(y_true, y_pred).mean())
    n_subsample = int(n_test * 0.9)
    scores = np.empty(n_bootstrap, dtype=np.float64)
    for i in range(n_bootstrap):
        boot_idx = rng.choice(n_test, size=n_subsample, replace=False)
        scores[i] = float(correlation_score(y_true[boot_idx], y_pred[boot_idx]).mean())
    return point, float(np.percentile(scores, 2.5)), float(np.percentile(scores, 97.5)), scores
