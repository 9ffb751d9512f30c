"""CPU oracle for the visreps RSA hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

This file is a CPU restatement (numpy / scipy / torch-CPU) of the arithmetic the
reference performs on the path named by BASELINE.json's north_star:

    compute_rdm              /root/reference/visreps/analysis/rsa.py:59-93
    _rank                    /root/reference/visreps/analysis/rsa.py:50-52
    compute_rdm_correlation  /root/reference/visreps/analysis/rsa.py:96-129
    _kendall_tau_a           /root/reference/visreps/analysis/rsa.py:22-40
    compute_rsa              /root/reference/visreps/analysis/rsa.py:132-281
    inlined bootstrap loops  /root/reference/visreps/evals.py:355-373, :506-522
    bootstrap_rdm_correlation
        /root/reference/experiments/reconstruction_analysis/run_reconstruction.py:137-158
    _concept_average_exact   /root/reference/visreps/analysis/rsa.py:284-305

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it.  The product package (visreps_b200/) never does; it fails
loudly when the CUDA library is missing.

Third-party arithmetic the reference delegates to, and that this oracle therefore
also delegates to (same installed versions: scipy 1.18.1, numpy 2.3.5, torch 2.11;
the reference's lock file pins scipy 1.16.2 / numpy 2.3.3 / torch 2.8.0):
    scipy.stats.spearmanr / pearsonr / kendalltau / rankdata,
    numpy.random.RandomState (MT19937 legacy stream), numpy.percentile (linear),
    torch CPU sgemm.

Parity pinning: oracle/gen_golden.py imports the UNMODIFIED reference module from
/root/reference (with a two-symbol stub for visreps.utils), runs it on seeded
inputs, asserts this oracle reproduces it, and writes tests/golden/*.npz.
tests/test_oracle_golden.py re-checks the oracle against those fixtures on any box.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import scipy.stats
import torch

SUBSAMPLE_FRACTION = 0.9  # rsa.py:234, evals.py:358


# --------------------------------------------------------------------------- #
# RDM construction
# --------------------------------------------------------------------------- #
def ordinal_row_ranks(x: torch.Tensor) -> torch.Tensor:
    """0-based ordinal rank of every entry within its row (rsa.py:50-52).

    The reference takes argsort twice; position in the first argsort's output
    is the rank.  Ties are split by sort position.
    """
    order = torch.argsort(x, dim=1)
    return torch.argsort(order, dim=1).float()


def rdm(x: torch.Tensor, correlation: str = "Pearson", correction: float = 1e-12) -> torch.Tensor:
    """Correlation-distance RDM, fp32, in the reference's operation order (rsa.py:59-93)."""
    kind = correlation.lower()
    if kind not in ("pearson", "spearman"):
        raise ValueError("correlation must be 'Pearson' or 'Spearman'")
    z = x.float().clone()
    if kind == "spearman":
        z = ordinal_row_ranks(z)
    d = z.size(1)
    z -= z.mean(dim=1, keepdim=True)
    sigma = z.pow(2).mean(dim=1).add(correction).sqrt()
    flat = sigma < correction * 10
    if flat.any():
        sigma[flat] = 1.0
    gram = (z @ z.T) / d
    r = gram / (sigma[:, None] * sigma[None, :] + correction)
    r.clamp_(-1, 1).fill_diagonal_(1.0)
    return 1.0 - r


def rdm_truth_f64(x: torch.Tensor, correction: float = 1e-12) -> np.ndarray:
    """Same formula evaluated in float64: the arbiter for the 1e-5 RDM budget."""
    z = x.detach().cpu().double().numpy().copy()
    d = z.shape[1]
    z -= z.mean(axis=1, keepdims=True)
    sigma = np.sqrt((z * z).mean(axis=1) + correction)
    sigma[sigma < correction * 10] = 1.0
    r = (z @ z.T) / d / (sigma[:, None] * sigma[None, :] + correction)
    np.clip(r, -1, 1, out=r)
    np.fill_diagonal(r, 1.0)
    return 1.0 - r


# --------------------------------------------------------------------------- #
# Upper triangle, ranks
# --------------------------------------------------------------------------- #
def triu_vector(m) -> np.ndarray:
    """Strict upper triangle in row-major order (rsa.py:111-113)."""
    a = m.detach().cpu().numpy() if isinstance(m, torch.Tensor) else np.asarray(m)
    iu = np.triu_indices(a.shape[0], k=1)
    return a[iu]


def rank2x(v: np.ndarray) -> np.ndarray:
    """2 * (tie-averaged 1-based rank) as int64.

    SciPy `rankdata(method='average')` semantics; evaluated on a float64 copy so
    the result is exact for any length (SciPy 1.18 returns
    float32 ranks for float32 input, which quantises above 2**23 elements).
    """
    r = scipy.stats.rankdata(np.asarray(v).astype(np.float64), method="average")
    return np.rint(2.0 * r).astype(np.int64)


def spearman_from_rank2x(r2a: np.ndarray, r2b: np.ndarray) -> float:
    """Pearson correlation of two rank vectors, integer-exact numerator/denominators."""
    n = r2a.shape[0]
    da = r2a.astype(np.int64) - (n + 1)
    db = r2b.astype(np.int64) - (n + 1)

    def exact_dot(u, v):
        # |u|,|v| < n, so a block of 2**10 products stays below 2**63 for n < 2**26;
        # block sums are then added as Python ints (no overflow at any n).
        total = 0
        for s in range(0, n, 1024):
            total += int(np.dot(u[s:s + 1024], v[s:s + 1024]))
        return total

    if n < 2_000_000:  # n**3 < 2**63: one int64 dot is exact
        sab, saa, sbb = int(np.dot(da, db)), int(np.dot(da, da)), int(np.dot(db, db))
    else:
        sab, saa, sbb = exact_dot(da, db), exact_dot(da, da), exact_dot(db, db)
    if saa == 0 or sbb == 0:
        return float("nan")
    return sab / math.sqrt(saa) / math.sqrt(sbb)


# --------------------------------------------------------------------------- #
# RDM-vs-RDM comparison
# --------------------------------------------------------------------------- #
def kendall_tau_a(x: np.ndarray, y: np.ndarray):
    """tau-a obtained from SciPy's tau-b (rsa.py:22-40)."""
    n = len(x)
    if n < 2:
        return float("nan"), float("nan")
    tau_b = scipy.stats.kendalltau(x, y).statistic
    if np.isnan(tau_b):
        return float("nan"), float("nan")
    n0 = n * (n - 1) // 2
    tx = sum(int(c) * (int(c) - 1) // 2 for c in np.unique(x, return_counts=True)[1])
    ty = sum(int(c) * (int(c) - 1) // 2 for c in np.unique(y, return_counts=True)[1])
    den = np.sqrt(np.float64(n0 - tx) * np.float64(n0 - ty))
    if den == 0:
        return float("nan"), float("nan")
    return float(tau_b * den / n0), float("nan")


_COMPARE = {
    "pearson": scipy.stats.pearsonr,
    "spearman": scipy.stats.spearmanr,
    "kendall": kendall_tau_a,
}


def rdm_compare(r1: torch.Tensor, r2: torch.Tensor, correlation: str = "Kendall") -> float:
    """Scalar agreement of two RDMs over the strict upper triangle (rsa.py:96-129)."""
    if r1.shape != r2.shape or r1.ndim != 2:
        raise ValueError("RDMs must share the same 2-D shape")
    n = r1.size(0)
    if n <= 1:
        return float("nan")
    v1, v2 = triu_vector(r1), triu_vector(r2)
    if v1.size == 0:
        return float("nan")
    kind = correlation.lower()
    if kind not in _COMPARE:
        raise ValueError("correlation must be 'Pearson', 'Spearman', or 'Kendall'")
    try:
        val = _COMPARE[kind](v1, v2)[0]
    except Exception:
        return float("nan")
    return float("nan") if np.isnan(val) else float(val)


# --------------------------------------------------------------------------- #
# Bootstrap
# --------------------------------------------------------------------------- #
def draw_subsamples(n: int, n_bootstrap: int, seed: int = 42,
                    rng: Optional[np.random.RandomState] = None) -> np.ndarray:
    """The reference's host RNG stream: B draws of int(0.9 n) stimuli without replacement.

    Order of draws and of indices inside a draw is the reference's
    (evals.py:356-364, rsa.py:249-251).
    """
    rng = np.random.RandomState(seed) if rng is None else rng
    m = int(n * SUBSAMPLE_FRACTION)
    out = np.empty((n_bootstrap, m), dtype=np.int64)
    for b in range(n_bootstrap):
        out[b] = rng.choice(n, size=m, replace=False)
    return out


def bootstrap_compare_with_indices(model_rdm: torch.Tensor, neural_rdm: torch.Tensor,
                                   method: str, index_sets: np.ndarray) -> np.ndarray:
    """One comparison per index set on the sub-indexed RDMs (evals.py:361-369)."""
    scores = np.empty(len(index_sets), dtype=np.float64)
    for b, idx in enumerate(index_sets):
        t = torch.from_numpy(np.asarray(idx, dtype=np.int64))
        scores[b] = rdm_compare(model_rdm[t][:, t], neural_rdm[t][:, t], correlation=method.capitalize())
    return scores


def bootstrap_compare(model_rdm: torch.Tensor, neural_rdm: torch.Tensor, method: str,
                      n_bootstrap: int = 1000, seed: int = 42):
    """(score, ci_low, ci_high, scores) as run_reconstruction.py:137-158 returns them."""
    point = rdm_compare(model_rdm, neural_rdm, correlation=method.capitalize())
    idx = draw_subsamples(neural_rdm.size(0), n_bootstrap, seed)
    scores = bootstrap_compare_with_indices(model_rdm, neural_rdm, method, idx)
    lo = float(np.percentile(scores, 2.5))
    hi = float(np.percentile(scores, 97.5))
    return point, lo, hi, scores.tolist()


# --------------------------------------------------------------------------- #
# Train/test RSA (layer selection + evaluation + bootstrap)
# --------------------------------------------------------------------------- #
@dataclass
class Split:
    """Field-compatible stand-in for AlignmentData (alignment.py:14-20)."""
    activations: Dict[str, torch.Tensor]
    neural: torch.Tensor
    stimulus_ids: Optional[List[str]] = None
    concept_image_ids: Optional[Dict[str, List[str]]] = None


def _flatten(a: torch.Tensor) -> torch.Tensor:
    return a.flatten(start_dim=1) if a.ndim > 2 else a


def rsa_traintest(cfg, selection, evaluation, n_select=None, bootstrap=True, n_bootstrap=1000,
                  seed=42, verbose=False, re_extract_fn: Optional[Callable] = None) -> List[Dict]:
    """Restatement of compute_rsa (rsa.py:132-281): one RandomState feeds both the
    optional selection subsample and, afterwards, the bootstrap draws."""
    method = cfg.get("compare_method", "spearman").lower()
    rng = np.random.RandomState(seed)
    n_train = selection.neural.size(0)
    n_test = evaluation.neural.size(0)
    if n_select is not None and n_select < n_train:
        pick = rng.choice(n_train, size=n_select, replace=False)
    else:
        pick = np.arange(n_train)

    neural_sel = rdm(selection.neural[pick])
    table, best, best_score = [], None, -float("inf")
    for name, acts in selection.activations.items():
        s = rdm_compare(rdm(_flatten(acts[pick])), neural_sel, correlation=method.capitalize())
        table.append({"layer": name, "score": s})
        if s > best_score:
            best, best_score = name, s

    if re_extract_fn is not None:
        exact, _ = re_extract_fn(best, evaluation.stimulus_ids)
        test_acts = _flatten(exact)
    else:
        test_acts = _flatten(evaluation.activations[best])
    neural_test = rdm(evaluation.neural)
    model_test = rdm(test_acts)
    point = rdm_compare(model_test, neural_test, correlation=method.capitalize())

    lo = hi = None
    scores = None
    if bootstrap:
        idx = draw_subsamples(n_test, n_bootstrap, rng=rng)
        s = bootstrap_compare_with_indices(model_test, neural_test, method, idx)
        lo, hi = float(np.percentile(s, 2.5)), float(np.percentile(s, 97.5))
        scores = s.tolist()
    out = {"layer": best, "compare_method": method, "score": point, "ci_low": lo, "ci_high": hi,
           "analysis": "rsa", "layer_selection_scores": table}
    if scores is not None:
        out["bootstrap_scores"] = scores
    return [out]


def concept_average(raw_acts: torch.Tensor, raw_ids: Sequence, data) -> torch.Tensor:
    """Per-concept mean of per-image activations in data.stimulus_ids order (rsa.py:284-305)."""
    where = {str(k): i for i, k in enumerate(raw_ids)}
    rows = []
    for concept in data.stimulus_ids:
        sel = [where[s] for s in data.concept_image_ids[concept] if s in where]
        rows.append(raw_acts[sel].float().mean(0) if sel else torch.zeros(raw_acts.size(1)))
    return torch.stack(rows).to(raw_acts.dtype)
