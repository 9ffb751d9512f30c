"""Seeded input cases shared by oracle/gen_golden.py and tests/ (test infrastructure).

Everything is drawn from numpy's legacy MT19937 RandomState so the same arrays are
rebuilt bit-identically on any box; only reference OUTPUTS are stored in tests/golden/.
"""
from __future__ import annotations

import numpy as np
import torch

# name -> (n stimuli, d features, seed)
RDM_CASES = {
    "n8_d5": (8, 5, 101),
    "n64_d48": (64, 48, 102),
    "n220_d96": (220, 96, 103),       # NSD-synthetic stimulus count (evals.py:442)
    "n130_d66": (130, 66, 104),       # THINGS behaviour width (d = 66), ragged n
    "n1000_d4096": (1000, 4096, 105), # NSD shared-1000 x fc7
}

# pairs compared / bootstrapped: name -> (model case, neural (n, V, seed), mixing weight)
PAIR_CASES = {
    "p64": ("n64_d48", (64, 30, 201), 0.7),
    "p220": ("n220_d96", (220, 120, 202), 0.5),
    "p1000": ("n1000_d4096", (1000, 300, 203), 0.3),
}
N_GOLDEN_REPLICATES = {"p64": 16, "p220": 16, "p1000": 8}


def features(n: int, d: int, seed: int) -> torch.Tensor:
    """Post-ReLU-like activations: non-negative, about half zeros."""
    rs = np.random.RandomState(seed)
    latent = rs.standard_normal((n, 16)).astype(np.float32)
    w = (rs.standard_normal((16, d)) / 4.0).astype(np.float32)
    x = latent @ w + rs.standard_normal((n, d)).astype(np.float32)
    return torch.from_numpy(np.maximum(x, 0.0).astype(np.float32))


def neural_like(model_x: torch.Tensor, n: int, v: int, seed: int, mix: float) -> torch.Tensor:
    """Voxel responses sharing structure with the model features so scores are non-trivial."""
    rs = np.random.RandomState(seed)
    k = min(model_x.shape[1], 32)
    proj = rs.standard_normal((k, v)).astype(np.float32)
    y = mix * (model_x[:, :k].numpy() @ proj) + rs.standard_normal((n, v)).astype(np.float32)
    return torch.from_numpy(y.astype(np.float32))


def rdm_input(name: str) -> torch.Tensor:
    n, d, seed = RDM_CASES[name]
    return features(n, d, seed)


def pair_inputs(name: str):
    model_case, (n, v, seed), mix = PAIR_CASES[name]
    x = rdm_input(model_case)
    return x, neural_like(x, n, v, seed, mix)


def tied_vectors(length: int, levels: int, seed: int):
    """Two float32 vectors with heavy ties (values on a coarse grid), incl. -0.0 / +0.0."""
    rs = np.random.RandomState(seed)
    a = (rs.randint(0, levels, size=length) / np.float32(levels)).astype(np.float32)
    b = (a + rs.randint(-2, 3, size=length) / np.float32(levels)).astype(np.float32)
    a[rs.randint(0, length, size=max(1, length // 50))] = np.float32(-0.0)
    return a, b


def vector_to_rdm(v: np.ndarray, n: int) -> torch.Tensor:
    """Symmetric matrix whose strict upper triangle (row-major) is v, zero diagonal."""
    m = np.zeros((n, n), dtype=np.float32)
    iu = np.triu_indices(n, k=1)
    m[iu] = v
    m.T[iu] = v
    return torch.from_numpy(m)


def alignment_case(seed: int = 42):
    """200 train / 50 test / 100 voxels, a 'good' and a 'bad' layer
    (shape of the fixture at /root/reference/tests/test_rsa_bootstrap.py:79-113)."""
    rs = np.random.RandomState(seed)
    g = lambda *s: torch.from_numpy(rs.standard_normal(s).astype(np.float32))
    n_train, n_test, v = 200, 50, 100
    neural_train, neural_test = g(n_train, v), g(n_test, v)
    train = dict(
        activations={"bad": g(n_train, v), "good": neural_train + 0.1 * g(n_train, v),
                     "conv": (neural_train + 0.5 * g(n_train, v)).reshape(n_train, 4, 5, 5)},
        neural=neural_train, stimulus_ids=[str(i) for i in range(n_train)])
    test = dict(
        activations={"bad": g(n_test, v), "good": neural_test + 0.1 * g(n_test, v),
                     "conv": (neural_test + 0.5 * g(n_test, v)).reshape(n_test, 4, 5, 5)},
        neural=neural_test, stimulus_ids=[str(i) for i in range(n_train, n_train + n_test)])
    return train, test
