"""The CPU oracle against the fixtures produced by the UNMODIFIED reference
(oracle/gen_golden.py).  This is what pins the oracle; it runs on any box."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cases, rsa_oracle as orc


@pytest.fixture(scope="module")
def rdm_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "rdm.npz"))


@pytest.fixture(scope="module")
def cmp_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "compare.npz"))


@pytest.mark.parametrize("name", list(cases.RDM_CASES))
def test_rdm_matches_reference(name, rdm_golden):
    x = cases.rdm_input(name)
    keep = x.clone()
    r = orc.rdm(x)
    assert torch.equal(x, keep), "input must not be mutated (rsa.py:76)"
    if f"{name}/pearson" in rdm_golden:
        np.testing.assert_array_equal(r.numpy(), rdm_golden[f"{name}/pearson"])
        np.testing.assert_array_equal(orc.rdm(x, correlation="Spearman").numpy(), rdm_golden[f"{name}/spearman"])
    else:
        np.testing.assert_array_equal(orc.triu_vector(r)[::97], rdm_golden[f"{name}/pearson_triu_stride97"])
    assert float(np.abs(r.numpy() - orc.rdm_truth_f64(x)).max()) < 1e-5
    assert r.dtype == torch.float32 and float(r.diagonal().abs().max()) == 0.0


def test_known_answer_rows(rdm_golden):
    simple = torch.tensor([[1., 2., 3.], [4., 5., 6.], [1., 2., 3.], [7., 8., 9.]])
    r = orc.rdm(simple)
    np.testing.assert_array_equal(r.numpy(), rdm_golden["simple/pearson"])
    assert abs(float(r[0, 2])) < 1e-4                       # identical rows -> 0
    anti = torch.tensor([[1., 2., 3., 4.], [4., 3., 2., 1.], [1., 1., 1., 1.], [2., 4., 6., 9.]])
    ra = orc.rdm(anti)
    np.testing.assert_array_equal(ra.numpy(), rdm_golden["anti_const/pearson"])
    assert abs(float(ra[0, 1]) - 2.0) < 1e-4                # anti-correlated rows -> 2
    assert torch.isfinite(ra).all()                         # constant row: std := 1


def test_bad_arguments():
    with pytest.raises(ValueError):
        orc.rdm(torch.randn(4, 3), correlation="cosine")
    with pytest.raises(ValueError):
        orc.rdm_compare(torch.zeros(5, 5), torch.zeros(6, 6))
    with pytest.raises(ValueError):
        orc.rdm_compare(torch.zeros(5, 5), torch.zeros(5, 5), correlation="cosine")
    assert np.isnan(orc.rdm_compare(torch.zeros(1, 1), torch.zeros(1, 1), correlation="Spearman"))
    assert np.isnan(orc.rdm_compare(torch.ones(6, 6), torch.ones(6, 6), correlation="Spearman"))


@pytest.mark.parametrize("name", list(cases.PAIR_CASES))
def test_compare_and_bootstrap_match_reference(name, cmp_golden):
    x, y = cases.pair_inputs(name)
    a, b = orc.rdm(x), orc.rdm(y)
    for method in ("spearman", "pearson", "kendall"):
        key = f"{name}/{method}"
        if key in cmp_golden:
            assert orc.rdm_compare(a, b, correlation=method.capitalize()) == float(cmp_golden[key])
    va, vb = orc.triu_vector(a), orc.triu_vector(b)
    ra, rb = orc.rank2x(va), orc.rank2x(vb)
    np.testing.assert_array_equal(ra[:64], cmp_golden[f"{name}/rank2x_a_head"])
    np.testing.assert_array_equal(rb[:64], cmp_golden[f"{name}/rank2x_b_head"])
    assert int(ra.sum()) == int(cmp_golden[f"{name}/rank2x_a_sum"])
    assert abs(orc.spearman_from_rank2x(ra, rb) - float(cmp_golden[f"{name}/spearman"])) < 1e-9
    nb = cases.N_GOLDEN_REPLICATES[name]
    idx = orc.draw_subsamples(a.shape[0], nb, 42)
    np.testing.assert_array_equal(idx, cmp_golden[f"{name}/boot_idx"])
    if a.shape[0] <= 256:   # the n=1000 replicates cost ~0.25 s each on the CPU: keep two
        reps = slice(None)
    else:
        reps = slice(0, 2)
    for method in ("spearman", "pearson"):
        got = orc.bootstrap_compare_with_indices(a, b, method, idx[reps])
        np.testing.assert_array_equal(got, cmp_golden[f"{name}/boot_{method}"][reps])


@pytest.mark.parametrize("tag,n,levels,seed", [("ties24", 24, 7, 301), ("ties90", 90, 40, 302)])
def test_tie_heavy_vectors(tag, n, levels, seed, cmp_golden):
    N = n * (n - 1) // 2
    va, vb = cases.tied_vectors(N, levels, seed)
    np.testing.assert_array_equal(orc.rank2x(va), cmp_golden[f"{tag}/rank2x_a"])
    np.testing.assert_array_equal(orc.rank2x(vb), cmp_golden[f"{tag}/rank2x_b"])
    a, b = cases.vector_to_rdm(va, n), cases.vector_to_rdm(vb, n)
    for method in ("spearman", "pearson", "kendall"):
        assert orc.rdm_compare(a, b, correlation=method.capitalize()) == float(cmp_golden[f"{tag}/{method}"])
    assert abs(orc.spearman_from_rank2x(orc.rank2x(va), orc.rank2x(vb)) - float(cmp_golden[f"{tag}/spearman"])) < 1e-12


def test_compute_rsa_matches_reference(golden_dir):
    with open(os.path.join(golden_dir, "compute_rsa.json")) as f:
        gold = json.load(f)["results"]
    tr, te = cases.alignment_case()
    for tag, kw in {"all_train": dict(n_select=None), "subsample": dict(n_select=120)}.items():
        got = orc.rsa_traintest({"compare_method": "spearman"}, orc.Split(**tr), orc.Split(**te), bootstrap=True,
                                n_bootstrap=25, seed=7, **kw)[0]
        assert got == gold[f"{tag}/spearman"]
        assert got["layer"] == "good" and len(got["bootstrap_scores"]) == 25


def test_kendall_hand_counts():
    # hand-counted tau-a examples (reference tests :383-394, :1124-1135)
    t, _ = orc.kendall_tau_a(np.array([1., 2., 3.]), np.array([1., 3., 2.]))
    assert t == pytest.approx(1 / 3)
    t, _ = orc.kendall_tau_a(np.array([1., 2., 3.]), np.array([3., 1., 2.]))
    assert t == pytest.approx(-1 / 3)
    assert np.isnan(orc.kendall_tau_a(np.array([1.]), np.array([1.]))[0])


def test_encoding_oracle_correlation_score_is_columnwise_pearson():
    """oracle/encoding_oracle.py restates himalaya's correlation_score (absent here): per column it must be
    scipy's Pearson r, and the bootstrap must consume the generator like the reference loop."""
    import scipy.stats
    from oracle import encoding_oracle as enc
    rs = np.random.RandomState(5)
    y = rs.standard_normal((60, 9)).astype(np.float32)
    p = (0.6 * y + rs.standard_normal((60, 9))).astype(np.float32)
    got = enc.correlation_score(y, p)
    want = [scipy.stats.pearsonr(y[:, v].astype(np.float64), p[:, v].astype(np.float64))[0] for v in range(9)]
    np.testing.assert_allclose(got, want, atol=2e-6)
    rng = np.random.RandomState(42)
    rng.permutation(200)                                       # encoding_score.py:131 consumes the generator first
    point, lo, hi, scores = enc.bootstrap_encoding(y, p, n_bootstrap=7, rng=rng)
    ref = np.random.RandomState(42)
    ref.permutation(200)
    idx0 = ref.choice(60, size=54, replace=False)
    assert scores[0] == float(enc.correlation_score(y[idx0], p[idx0]).mean())
    assert lo <= np.median(scores) <= hi and abs(point - np.mean(want)) < 1e-5
