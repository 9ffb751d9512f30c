"""Device versions of the scoring blocks that `visreps/evals.py` inlines around the RSA numerics.

The reference evaluates NSD/TVSD in three places that all run the same loops over (region, subject):

  * phase 1, layer selection (evals.py:249-287): for every pair, up to `n_select` training stimuli drawn by a
    fresh `RandomState(42)`, one Pearson RDM per layer, one comparison per layer, first maximum wins;
  * the scoring block of `_eval_rsa` (evals.py:333-398) and of `_eval_rsa_nsd_synthetic` (evals.py:489-546):
    per pair a neural RDM, the point estimate against the pair's best layer and, optionally, 1000 90 %-subsample
    replicates from a `RandomState(42)` that is re-created for every pair -- so all pairs see the same index sets.

Here each block is one grouped device job (`rsa_grouped`): all layers x the pair's neural matrix for the
selection, all distinct best layers x all pairs for the scoring, with the rankings shared across pairs.  The
result dictionaries have the reference's keys and value types, in the reference's (region, subject) order.
Model loading, feature extraction, the neural data loaders and `save_results` stay where they are.
"""
from __future__ import annotations

from typing import Dict, Hashable, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import rsa

SELECTION_SEED = 42      # evals.py:259 and :356 hard-code RandomState(42)


def selection_indices(n_train: int, n_select: Optional[int]) -> np.ndarray:
    """evals.py:257-262: `RandomState(42).choice(n_train, n_select, replace=False)` when the pair has more
    training stimuli than `n_select`, else all of them in order."""
    if n_select is not None and n_select < n_train:
        return np.random.RandomState(SELECTION_SEED).choice(n_train, size=n_select, replace=False)
    return np.arange(n_train)


def select_layer(train_acts: Mapping[str, torch.Tensor], train_neural: torch.Tensor, method: str = "spearman",
                 n_select: Optional[int] = 1000) -> Tuple[Optional[str], List[Dict]]:
    """Phase-1 layer selection for one (region, subject) pair (evals.py:251-276).

    train_acts: {layer: (n_train, ...)} activations aligned with train_neural (n_train, V).
    Returns (best_layer, [{"layer": name, "score": float}, ...]) with the reference's tie rule: the first
    layer whose score is strictly greater than every earlier one (a NaN score never wins)."""
    n_train = int(train_neural.size(0))
    sel_idx = selection_indices(n_train, n_select)
    idx_t = torch.from_numpy(sel_idx)
    names = list(train_acts.keys())
    feats = [rsa._flat(train_acts[k][idx_t.to(train_acts[k].device)]) for k in names]
    res = rsa.rsa_grouped(feats, [train_neural[idx_t.to(train_neural.device)]], method, bootstrap=False)
    scores, best_layer, best_score = [], None, -float("inf")
    for name, s in zip(names, res["point"][:, 0]):
        s = float(s)
        scores.append({"layer": name, "score": s})
        if s > best_score:
            best_score, best_layer = s, name
    return best_layer, scores


def score_pairs(model_features: Mapping[str, torch.Tensor], neural_test: Mapping[Hashable, torch.Tensor],
                best_layers: Mapping[Hashable, str], method: str = "spearman", bootstrap: bool = False,
                n_bootstrap: int = 1000,
                selection_scores: Optional[Mapping[Hashable, Sequence[Dict]]] = None) -> List[Dict]:
    """The per-(region, subject) scoring block of `_eval_rsa` / `_eval_rsa_nsd_synthetic`
    (evals.py:333-398, :489-546) for all pairs in one device job.

    model_features: {layer: (n_test, ...)} exact activations of every distinct best layer (host or device);
    neural_test:    {pair key: (n_test, V)} test responses, iterated in the caller's (region, subject) order;
    best_layers:    {pair key: layer name}.
    Returns one dict per pair, in `neural_test` order, with the keys of evals.py:375-386:
    layer, compare_method, score, ci_low, ci_high, analysis, layer_selection_scores[, bootstrap_scores].
    """
    method = method.lower()
    keys = list(neural_test.keys())
    layers = sorted({best_layers[k] for k in keys})            # the reference re-extracts sorted(unique layers), :309
    col = {name: i for i, name in enumerate(layers)}
    feats = [rsa._flat(model_features[name]) for name in layers]
    neur = [torch.as_tensor(neural_test[k], dtype=torch.float32) for k in keys]
    res = rsa.rsa_grouped(feats, neur, method, bootstrap=bootstrap, n_bootstrap=n_bootstrap, seed=SELECTION_SEED)
    out = []
    for j, k in enumerate(keys):
        i = col[best_layers[k]]
        result = {"layer": best_layers[k], "compare_method": method, "score": float(res["point"][i, j]),
                  "ci_low": None, "ci_high": None, "analysis": "rsa",
                  "layer_selection_scores": list(selection_scores[k]) if selection_scores is not None else []}
        if bootstrap:
            result["ci_low"] = float(res["ci_low"][i, j])
            result["ci_high"] = float(res["ci_high"][i, j])
            result["bootstrap_scores"] = res["scores"][i, j].tolist()
        out.append(result)
    return out


def eval_rsa_scores(train_acts_by_pair: Mapping[Hashable, Mapping[str, torch.Tensor]],
                    train_neural_by_pair: Mapping[Hashable, torch.Tensor],
                    test_features: Mapping[str, torch.Tensor], test_neural_by_pair: Mapping[Hashable, torch.Tensor],
                    method: str = "spearman", n_select: Optional[int] = 1000, bootstrap: bool = False,
                    n_bootstrap: int = 1000) -> List[Dict]:
    """Phase 1 + scoring of `_eval_rsa` (evals.py:249-287 + :333-398) from already extracted activations:
    per-pair layer selection on the training split, then every pair's best layer scored on the test split.
    `test_features` must hold the exact test activations of every layer that can be selected."""
    best, sel_scores = {}, {}
    for k in test_neural_by_pair.keys():
        best[k], sel_scores[k] = select_layer(train_acts_by_pair[k], train_neural_by_pair[k], method, n_select)
    return score_pairs(test_features, test_neural_by_pair, best, method, bootstrap, n_bootstrap, sel_scores)


def _compare_rdms(rdm1: torch.Tensor, rdm2: torch.Tensor, correlation: str) -> float:
    """`experiments/binary_pc_rsa/main.py:28-38`: the same comparison with a positional method name."""
    return rsa.compute_rdm_correlation(rdm1, rdm2, correlation=correlation)
