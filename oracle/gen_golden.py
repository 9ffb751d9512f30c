"""Generate tests/golden/*.npz by running the UNMODIFIED reference (test infrastructure).

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py

It imports /root/reference/visreps/analysis/rsa.py as-is.  That module does
`from visreps.utils import console, rprint` (rsa.py:10) and visreps/utils.py imports
omegaconf, which is not installed here, so `visreps.utils` is pre-seeded with a
two-symbol stub (neither symbol is on the numeric path).  Every reference output is
also compared with oracle/rsa_oracle.py on the spot: the script fails if the oracle
does not reproduce the reference, which is what pins the oracle.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REFERENCE = os.environ.get("VISREPS_REFERENCE", "/root/reference")


def import_reference():
    stub = types.ModuleType("visreps.utils")
    from rich.console import Console
    stub.console = Console(quiet=True)
    stub.rprint = lambda *a, **k: None
    pkg = types.ModuleType("visreps")
    pkg.__path__ = [os.path.join(REFERENCE, "visreps")]
    ana = types.ModuleType("visreps.analysis")
    ana.__path__ = [os.path.join(REFERENCE, "visreps", "analysis")]
    sys.modules.update({"visreps": pkg, "visreps.utils": stub, "visreps.analysis": ana})
    import importlib
    return importlib.import_module("visreps.analysis.rsa")


def main():
    from oracle import cases, rsa_oracle as orc
    ref = import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    meta = {"reference": REFERENCE, "torch": torch.__version__, "numpy": np.__version__}
    import scipy
    meta["scipy"] = scipy.__version__

    # ---- RDMs ---------------------------------------------------------------
    blob = {}
    for name in cases.RDM_CASES:
        x = cases.rdm_input(name)
        r = ref.compute_rdm(x)
        o = orc.rdm(x)
        assert torch.equal(r, o), f"oracle rdm != reference for {name}"
        truth = orc.rdm_truth_f64(x)
        n = x.shape[0]
        if n <= 256:
            blob[f"{name}/pearson"] = r.numpy()
            rs = ref.compute_rdm(x, correlation="Spearman")
            assert torch.equal(rs, orc.rdm(x, correlation="Spearman"))
            blob[f"{name}/spearman"] = rs.numpy()
        else:  # too big for a fixture: keep a strided sample of the upper triangle
            v = orc.triu_vector(r)
            blob[f"{name}/pearson_triu_stride97"] = v[::97].copy()
        blob[f"{name}/max_abs_vs_f64"] = np.float64(np.abs(r.numpy().astype(np.float64) - truth).max())
        blob[f"{name}/bit_symmetric"] = np.bool_(torch.equal(r, r.T))
    # known-answer rows (tests/test_rsa_bootstrap.py:65-75, :145-150, :214-225)
    simple = torch.tensor([[1., 2., 3.], [4., 5., 6.], [1., 2., 3.], [7., 8., 9.]])
    blob["simple/pearson"] = ref.compute_rdm(simple).numpy()
    anti = torch.tensor([[1., 2., 3., 4.], [4., 3., 2., 1.], [1., 1., 1., 1.], [2., 4., 6., 9.]])
    blob["anti_const/pearson"] = ref.compute_rdm(anti).numpy()
    assert torch.equal(ref.compute_rdm(anti), orc.rdm(anti))
    np.savez_compressed(os.path.join(out_dir, "rdm.npz"), **blob)

    # ---- comparisons + bootstrap -------------------------------------------
    blob = {}
    for name in cases.PAIR_CASES:
        x, y = cases.pair_inputs(name)
        a, b = ref.compute_rdm(x), ref.compute_rdm(y)
        va, vb = orc.triu_vector(a), orc.triu_vector(b)
        for method in ("Spearman", "Pearson") + (("Kendall",) if x.shape[0] <= 256 else ()):
            s = ref.compute_rdm_correlation(a, b, correlation=method)
            assert s == orc.rdm_compare(a, b, correlation=method), (name, method)
            blob[f"{name}/{method.lower()}"] = np.float64(s)
        # rank adjudicator + integer-exact Spearman agree with SciPy's float path
        ra, rb = orc.rank2x(va), orc.rank2x(vb)
        exact = orc.spearman_from_rank2x(ra, rb)
        assert abs(exact - float(blob[f"{name}/spearman"])) < 1e-9
        blob[f"{name}/rank2x_a_sum"] = np.int64(ra.sum())
        blob[f"{name}/rank2x_a_head"] = ra[:64].copy()
        blob[f"{name}/rank2x_b_head"] = rb[:64].copy()
        blob[f"{name}/n_distinct_a"] = np.int64(np.unique(va).size)
        # bootstrap: the reference loop, evals.py:356-369, restated on the imported functions
        nb = cases.N_GOLDEN_REPLICATES[name]
        rng = np.random.RandomState(42)
        n = a.size(0)
        m = int(n * 0.9)
        idx_sets = np.empty((nb, m), dtype=np.int64)
        for method in ("Spearman", "Pearson"):
            rng = np.random.RandomState(42)
            sc = np.empty(nb)
            for i in range(nb):
                idx_np = rng.choice(n, size=m, replace=False)
                idx_sets[i] = idx_np
                idx = torch.from_numpy(idx_np)
                sc[i] = ref.compute_rdm_correlation(a[idx][:, idx], b[idx][:, idx], correlation=method)
            mine = orc.bootstrap_compare_with_indices(a, b, method.lower(), orc.draw_subsamples(n, nb, 42))
            assert np.array_equal(sc, mine), (name, method)
            blob[f"{name}/boot_{method.lower()}"] = sc
        blob[f"{name}/boot_idx"] = idx_sets.astype(np.int32)
    # tie-heavy vectors (incl. -0.0/+0.0): reference path = scipy on the triu of a symmetric matrix
    for tag, (n, levels, seed) in {"ties24": (24, 7, 301), "ties90": (90, 40, 302)}.items():
        N = n * (n - 1) // 2
        va, vb = cases.tied_vectors(N, levels, seed)
        a, b = cases.vector_to_rdm(va, n), cases.vector_to_rdm(vb, n)
        for method in ("Spearman", "Pearson", "Kendall"):
            s = ref.compute_rdm_correlation(a, b, correlation=method)
            assert s == orc.rdm_compare(a, b, correlation=method)
            blob[f"{tag}/{method.lower()}"] = np.float64(s)
        blob[f"{tag}/rank2x_a"] = orc.rank2x(va)
        blob[f"{tag}/rank2x_b"] = orc.rank2x(vb)
    np.savez_compressed(os.path.join(out_dir, "compare.npz"), **blob)

    # ---- compute_rsa (layer selection + bootstrap) ---------------------------
    from dataclasses import dataclass

    @dataclass
    class AD:
        activations: dict
        neural: torch.Tensor
        stimulus_ids: list = None
        concept_image_ids: dict = None

    tr, te = cases.alignment_case()
    results = {}
    for tag, kw in {"all_train": dict(n_select=None), "subsample": dict(n_select=120)}.items():
        for method in ("spearman",):
            got = ref.compute_rsa({"compare_method": method}, AD(**tr), AD(**te), bootstrap=True,
                                  n_bootstrap=25, seed=7, **kw)
            mine = orc.rsa_traintest({"compare_method": method}, orc.Split(**tr), orc.Split(**te),
                                     bootstrap=True, n_bootstrap=25, seed=7, **kw)
            assert got == mine, tag
            results[f"{tag}/{method}"] = got[0]
    with open(os.path.join(out_dir, "compute_rsa.json"), "w") as f:
        json.dump({"meta": meta, "results": results}, f, indent=1)
    print("golden fixtures written to", out_dir)
    for fn in sorted(os.listdir(out_dir)):
        print(f"  {fn}: {os.path.getsize(os.path.join(out_dir, fn)) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
