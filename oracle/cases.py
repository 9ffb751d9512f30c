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


# --------------------------------------------------------------------------------------------------
# Synthetic `_eval_rsa` job (visreps/evals.py:222-398): what eval() hands to the function, plus a fake producer.
# --------------------------------------------------------------------------------------------------
class EvalCfg(dict):
    """Mapping with attribute access, standing in for the OmegaConf node `_eval_rsa` receives."""
    __getattr__ = dict.get


def eval_rsa_job(seed: int = 77, n_shared: int = 60, n_private: int = 150, bootstrap: bool = True, n_bootstrap: int = 12,
                 method: str = "spearman", n_select: int = 100):
    """2 regions x 2 subjects; every subject saw its own `n_private` training stimuli, all saw the `n_shared` test
    stimuli (the NSD design).  Three layers, one of them 4-D; SRP-like 'acts' for every stimulus and 'exact' activations
    (a different tensor, as the reference's re-extraction without SRP gives) for the test stimuli.
    Returns (cfg, model, acts, ids, all_data, subjects, regions)."""
    rs = np.random.RandomState(seed)
    subjects, regions = [0, 1], ["V1", "ventral"]
    shared = [f"shared{k:03d}" for k in range(n_shared)]
    private = {s: [f"s{s}_{k:03d}" for k in range(n_private)] for s in subjects}
    ids = []
    for k in range(max(n_shared, n_private)):                       # interleave so alignment has to pick rows
        if k < n_shared:
            ids.append(shared[k])
        for s in subjects:
            if k < n_private:
                ids.append(private[s][k])
    n_all = len(ids)
    latent = rs.standard_normal((n_all, 10)).astype(np.float32)
    g = lambda *shape: rs.standard_normal(shape).astype(np.float32)
    dims = {"conv1_post": (6, 4, 4), "conv3_post": (48,), "fc2_post": (32,)}
    exact, acts = {}, {}
    for li, (name, shp) in enumerate(dims.items()):
        d = int(np.prod(shp))
        w = g(10, d) * (0.2 + 0.4 * li)
        full = np.maximum(latent @ w + g(n_all, d), 0).astype(np.float32)
        exact[name] = torch.from_numpy(full.reshape((n_all,) + shp))
        proj = g(d, min(d, 24)) / np.sqrt(d)                           # stand-in for the SRP projection
        acts[name] = torch.from_numpy((full @ proj).astype(np.float32))
    row = {sid: i for i, sid in enumerate(ids)}
    neural = {}
    for ri, region in enumerate(regions):
        v = 20 + 15 * ri
        wv = g(10, v)
        neural[region] = {}
        for s in subjects:
            resp = (0.4 + 0.3 * ri) * (latent @ wv) + g(n_all, v) + 0.1 * s
            # responses are stored with a leading singleton axis, which np.stack(...).squeeze() removes (evals.py:343)
            neural[region][s] = {"train": {sid: resp[row[sid]] for sid in private[s]},
                                 "test": {sid: resp[row[sid]][None, :] for sid in shared}}
    stimuli = {sid: ("image", sid) for sid in ids}
    all_data = {"neural": neural, "shared_test_ids": list(shared), "stimuli": stimuli}
    model = {"exact": exact, "row": row}                               # what the fake extract_single_layer reads
    cfg = EvalCfg(compare_method=method, bootstrap=bootstrap, n_bootstrap=n_bootstrap, n_select=n_select, neural_dataset="nsd",
                  batchsize=32, num_workers=0, log_expdata=True, seed=1, reconstruct_from_pcs=False, pca_k=1)
    return cfg, model, acts, ids, all_data, subjects, regions


def fake_extract_single_layer(model, dataloader, device, layer_name, stimulus_ids=None):
    """Stand-in for visreps/models/utils.py:350-404: exact activations of one layer, flattened, CPU fp32, rows in
    `stimulus_ids` order (the dataloader is the dict of stimuli the fake `_make_loader` returned)."""
    keys = list(dataloader.keys())
    x = model["exact"][layer_name]
    flat = x.reshape(x.shape[0], -1)
    order = [k for k in (stimulus_ids if stimulus_ids is not None else keys) if k in model["row"] and k in dataloader]
    out = flat[[model["row"][k] for k in order]].float()
    return out, order


def results_cfg():
    """Identity fields of a run as `save_results` reads them (visreps/utils.py:300-311, :395-430)."""
    return dict(seed=1, epoch=20, neural_dataset="nsd", cfg_id=3, pca_labels=True, pca_n_classes=8, pca_labels_folder="pca_labels_alexnet",
                model_name="CustomCNN", checkpoint_dir="model_checkpoints/x", analysis="rsa", compare_method="spearman",
                reconstruct_from_pcs=False, pca_k=1, log_expdata=True)
