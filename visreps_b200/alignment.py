"""Boundary type and thin dispatch of the reference alignment layer.

Mirrors /root/reference/visreps/analysis/alignment.py: AlignmentData (:14-20),
_align_stimulus_level (:23-39), compute_traintest_alignment (:74-114, RSA branch only).
The encoding-score branch is a different analysis and is not part of this path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from .rsa import compute_rsa


@dataclass
class AlignmentData:
    """Bundled activations and neural data for one split (train or test)."""
    activations: Dict[str, torch.Tensor]           # {layer_name: (n_stimuli, features...)}
    neural: torch.Tensor                           # (n_stimuli, n_voxels)
    stimulus_ids: Optional[List[str]] = None       # ordered IDs matching rows
    concept_image_ids: Optional[Dict[str, List[str]]] = None  # THINGS: concept -> [image ids]


def _align_stimulus_level(acts_raw, targets, keys):
    """Keep the stimuli that have a neural target, in `keys` order.
    Returns (acts, neural, matched_ids)."""
    idx = [i for i, k in enumerate(keys) if str(k) in targets]
    matched = [str(keys[i]) for i in idx]
    if not matched:
        return {l: a[:0] for l, a in acts_raw.items()}, torch.empty(0, dtype=torch.float32), matched
    neural = torch.as_tensor(np.stack([targets[s] for s in matched]), dtype=torch.float32)
    return {l: a[idx] for l, a in acts_raw.items()}, neural, matched


def compute_traintest_alignment(cfg, train: AlignmentData, test: AlignmentData, verbose: bool = False,
                                re_extract_fn=None) -> List[dict]:
    """Dispatch on cfg['analysis']; only 'rsa' is served by this package."""
    analysis = cfg.get("analysis", "rsa").lower()
    if analysis == "encoding_score" and cfg.get("neural_dataset", "").lower() == "things-behavior":
        raise ValueError("Encoding score is not supported for things-behavior (behavioral embeddings "
                         "have no voxels to predict). Use analysis=rsa instead.")     # alignment.py:91-95
    if analysis == "rsa":
        return compute_rsa(cfg, train, test, n_select=cfg.get("n_select", None), bootstrap=cfg.get("bootstrap", True),
                           n_bootstrap=cfg.get("n_bootstrap", 1000), verbose=verbose, re_extract_fn=re_extract_fn)
    if analysis == "encoding_score":
        raise NotImplementedError("analysis=encoding_score is outside the RSA hot path this package accelerates")
    raise ValueError(f"Unknown analysis method: {analysis}")
