"""Behavioural checks in the spirit of the reference's own RSA test file
(/root/reference/tests/test_rsa_bootstrap.py: TestComputeRDM :120-231, TestComputeRDMCorrelation :238-343,
TestRDMMathCorrectness :907-1042, TestComputeRSA :427-651, TestEvalRSABootstrap :658-770,
TestRSAPipelineInvariants :1259-1479, TestEdgeCases :2024-2152), run against the CUDA path through the
drop-in Python surface.  SciPy is the in-test oracle exactly as in the reference's tests."""
import math

import numpy as np
import pytest
import scipy.stats
import torch

import visreps_b200 as vb
from visreps_b200 import rsa as vrsa

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


def _split(rs, n_train=200, n_test=50, v=100):
    g = lambda *s: torch.from_numpy(rs.standard_normal(s).astype(np.float32))
    ntr, nte = g(n_train, v), g(n_test, v)
    tr = vb.AlignmentData({"good": ntr + 0.1 * g(n_train, v), "bad": g(n_train, v)}, ntr, [str(i) for i in range(n_train)])
    te = vb.AlignmentData({"good": nte + 0.1 * g(n_test, v), "bad": g(n_test, v)}, nte,
                          [str(i) for i in range(n_train, n_train + n_test)])
    return tr, te


# ----------------------------------------------------------------------------- compute_rdm
def test_rdm_shape_symmetry_diagonal_range(dev):
    x = torch.randn(37, 19)
    r = vb.compute_rdm(x.to(dev)).cpu()
    assert r.shape == (37, 37) and r.dtype == torch.float32
    assert torch.equal(r, r.T) and torch.all(r.diagonal() == 0)
    assert float(r.min()) >= 0 and float(r.max()) <= 2


def test_rdm_entries_equal_one_minus_pairwise_pearson(dev):
    rs = np.random.RandomState(1)
    x = rs.standard_normal((12, 40)).astype(np.float32)
    r = vb.compute_rdm(torch.from_numpy(x).to(dev)).cpu().numpy()
    for i in range(12):
        for j in range(i + 1, 12):
            assert r[i, j] == pytest.approx(1 - scipy.stats.pearsonr(x[i], x[j])[0], abs=1e-5)


def test_spearman_rdm_equals_one_minus_pairwise_spearman(dev):
    rs = np.random.RandomState(2)
    x = rs.standard_normal((9, 60)).astype(np.float32)
    r = vb.compute_rdm(torch.from_numpy(x).to(dev), correlation="Spearman").cpu().numpy()
    for i in range(9):
        for j in range(i + 1, 9):
            assert r[i, j] == pytest.approx(1 - scipy.stats.spearmanr(x[i], x[j])[0], abs=1e-5)


def test_rdm_known_patterns(dev):
    base = torch.arange(1.0, 9.0)
    x = torch.stack([base, 3 * base + 2, -base, base.flip(0), torch.ones(8) * 5])
    r = vb.compute_rdm(x.to(dev)).cpu()
    assert float(r[0, 1]) == pytest.approx(0.0, abs=1e-4)      # affine copies: zero distance
    assert float(r[0, 2]) == pytest.approx(2.0, abs=1e-4)      # anti-correlated: distance two
    assert float(r[0, 3]) == pytest.approx(2.0, abs=1e-4)
    assert torch.isfinite(r).all()                             # constant row does not produce NaN/inf


def test_rdm_is_permutation_equivariant_and_leaves_input_alone(dev):
    x = torch.randn(25, 33)
    keep = x.clone()
    r = vb.compute_rdm(x.to(dev)).cpu()
    perm = torch.randperm(25)
    rp = vb.compute_rdm(x[perm].to(dev)).cpu()
    assert torch.equal(x, keep)
    assert float((rp - r[perm][:, perm]).abs().max()) <= 2e-6
    # scaling and shifting the features of a stimulus does not change its correlations
    r2 = vb.compute_rdm((x * 7.5 + 3).to(dev)).cpu()
    assert float((r2 - r).abs().max()) <= 1e-5


# ----------------------------------------------------------------------------- compute_rdm_correlation
@pytest.mark.parametrize("method", ["Spearman", "Pearson", "Kendall"])
def test_compare_values_against_scipy(method, dev):
    rs = np.random.RandomState(3)
    a = vb.compute_rdm(torch.from_numpy(rs.standard_normal((15, 8)).astype(np.float32)).to(dev))
    b = vb.compute_rdm(torch.from_numpy(rs.standard_normal((15, 8)).astype(np.float32)).to(dev))
    iu = np.triu_indices(15, 1)
    va, vbv = a.cpu().numpy()[iu], b.cpu().numpy()[iu]
    if method == "Kendall":
        tb = scipy.stats.kendalltau(va, vbv).statistic      # no ties here: tau-a == tau-b
        want = tb
    else:
        want = {"Spearman": scipy.stats.spearmanr, "Pearson": scipy.stats.pearsonr}[method](va, vbv)[0]
    assert vb.compute_rdm_correlation(a, b, correlation=method) == pytest.approx(want, abs=1e-5)
    assert vb.compute_rdm_correlation(a, a, correlation=method) == pytest.approx(1.0, abs=1e-6)
    assert vb.compute_rdm_correlation(a, b, correlation=method.upper()) == pytest.approx(want, abs=1e-5)  # case-insensitive


def test_compare_unrelated_and_inverted_rdms(dev):
    a = vb.compute_rdm(torch.randn(30, 10).to(dev))
    b = vb.compute_rdm(torch.randn(30, 10).to(dev))
    assert abs(vb.compute_rdm_correlation(a, b, correlation="Spearman")) < 0.3
    inv = 2.0 - a
    inv.fill_diagonal_(0.0)
    assert vb.compute_rdm_correlation(a, inv, correlation="Spearman") < -0.99
    # only the strict upper triangle is read: garbage on the diagonal / lower triangle changes nothing
    junk = a.clone()
    junk += torch.tril(torch.full_like(a, 9.0))
    assert vb.compute_rdm_correlation(junk, b, correlation="Pearson") == pytest.approx(
        vb.compute_rdm_correlation(a, b, correlation="Pearson"), abs=1e-12)


def test_compare_argument_errors(dev):
    with pytest.raises(ValueError):
        vb.compute_rdm_correlation(torch.zeros(5, 5), torch.zeros(6, 6))
    with pytest.raises(ValueError):
        vb.compute_rdm_correlation(torch.zeros(5), torch.zeros(5))
    with pytest.raises(ValueError):
        vb.compute_rdm_correlation(torch.eye(4), torch.eye(4), correlation="cosine")
    assert math.isnan(vb.compute_rdm_correlation(torch.zeros(1, 1), torch.zeros(1, 1)))


# ----------------------------------------------------------------------------- compute_rsa
def test_rsa_selects_the_aligned_layer_and_reports_a_ci(dev):
    tr, te = _split(np.random.RandomState(42))
    res = vb.compute_rsa({"compare_method": "spearman"}, tr, te, bootstrap=True, n_bootstrap=50, seed=42)
    assert isinstance(res, list) and len(res) == 1
    r = res[0]
    assert set(r) == {"layer", "compare_method", "score", "ci_low", "ci_high", "analysis", "layer_selection_scores", "bootstrap_scores"}
    assert r["layer"] == "good" and r["analysis"] == "rsa" and r["compare_method"] == "spearman"
    assert r["score"] > 0.5 and r["ci_low"] <= r["ci_high"]
    assert len(r["bootstrap_scores"]) == 50 and all(isinstance(s, float) for s in r["bootstrap_scores"])
    assert r["ci_low"] - 0.05 < r["score"] < r["ci_high"] + 0.05
    sel = {d["layer"]: d["score"] for d in r["layer_selection_scores"]}
    assert sel["good"] > sel["bad"] and list(sel) == ["good", "bad"]


def test_rsa_is_deterministic_per_seed_and_seed_sensitive(dev):
    tr, te = _split(np.random.RandomState(5))
    a = vb.compute_rsa({"compare_method": "spearman"}, tr, te, n_bootstrap=20, seed=1)[0]
    b = vb.compute_rsa({"compare_method": "spearman"}, tr, te, n_bootstrap=20, seed=1)[0]
    c = vb.compute_rsa({"compare_method": "spearman"}, tr, te, n_bootstrap=20, seed=2)[0]
    assert a == b
    assert a["score"] == c["score"] and a["bootstrap_scores"] != c["bootstrap_scores"]


def test_rsa_selection_does_not_see_the_evaluation_split(dev):
    rs = np.random.RandomState(6)
    tr, te = _split(rs)
    _, te2 = _split(np.random.RandomState(7))
    a = vb.compute_rsa({"compare_method": "spearman"}, tr, te, bootstrap=False)[0]
    b = vb.compute_rsa({"compare_method": "spearman"}, tr, te2, bootstrap=False)[0]
    assert a["layer_selection_scores"] == b["layer_selection_scores"] and a["score"] != b["score"]


def test_rsa_leaves_inputs_untouched_and_accepts_n_select(dev):
    tr, te = _split(np.random.RandomState(8))
    snap = {k: v.clone() for k, v in list(tr.activations.items()) + [("__n", tr.neural), ("__t", te.neural)]}
    full = vb.compute_rsa({"compare_method": "spearman"}, tr, te, n_bootstrap=5)[0]
    sub = vb.compute_rsa({"compare_method": "spearman"}, tr, te, n_select=80, n_bootstrap=5)[0]
    assert torch.equal(tr.neural, snap["__n"]) and torch.equal(te.neural, snap["__t"])
    assert all(torch.equal(tr.activations[k], snap[k]) for k in tr.activations)
    assert sub["layer"] == full["layer"] and sub["layer_selection_scores"] != full["layer_selection_scores"]
    assert sub["score"] == full["score"]          # evaluation does not depend on the selection subsample


def test_rsa_noise_lowers_the_score(dev):
    rs = np.random.RandomState(9)
    neural = torch.from_numpy(rs.standard_normal((60, 80)).astype(np.float32))
    scores = []
    for noise in (0.05, 0.5, 2.0, 8.0):
        acts = neural + noise * torch.from_numpy(rs.standard_normal((60, 80)).astype(np.float32))
        scores.append(vb.compute_rdm_correlation(vb.compute_rdm(acts.to(dev)), vb.compute_rdm(neural.to(dev)), correlation="Spearman"))
    assert scores == sorted(scores, reverse=True) and scores[0] > 0.9 > 0.3 > scores[-1]


# ----------------------------------------------------------------------------- the inlined bootstrap of _eval_rsa
def test_eval_rsa_style_bootstrap(dev):
    """evals.py:355-373: RandomState(42) per (region, subject), int(0.9 n) stimuli without replacement."""
    rs = np.random.RandomState(10)
    n = 73
    model = vb.compute_rdm(torch.from_numpy(rs.standard_normal((n, 50)).astype(np.float32)).to(dev))
    neural = vb.compute_rdm(torch.from_numpy(rs.standard_normal((n, 31)).astype(np.float32)).to(dev))
    idx = vrsa.draw_index_sets(n, 30, 42)
    assert idx.shape == (30, int(n * 0.9)) and all(len(set(r.tolist())) == idx.shape[1] for r in idx)
    point, lo, hi, scores = vb.bootstrap_rdm_correlation(model, neural, "spearman", n_bootstrap=30, seed=42)
    # the reference loop, literally, with the drop-in compute_rdm_correlation
    rng = np.random.RandomState(42)
    want = []
    for _ in range(30):
        t = torch.from_numpy(rng.choice(n, size=int(n * 0.9), replace=False)).to(dev)
        want.append(vb.compute_rdm_correlation(model[t][:, t], neural[t][:, t], correlation="Spearman"))
    np.testing.assert_allclose(scores, want, atol=1e-12, rtol=0)
    assert lo == pytest.approx(np.percentile(want, 2.5)) and hi == pytest.approx(np.percentile(want, 97.5))
    assert point == pytest.approx(vb.compute_rdm_correlation(model, neural, correlation="Spearman"), abs=1e-15)
    # every (region, subject) pair of a job sees the same index sets: grouped == pair by pair
    neural2 = vb.compute_rdm(torch.from_numpy(rs.standard_normal((n, 12)).astype(np.float32)).to(dev))
    grouped = vb.bootstrap_rdm_correlation_grouped([model], [neural, neural2], "spearman", n_bootstrap=30, seed=42)
    np.testing.assert_allclose(grouped["scores"][0, 0], scores, atol=1e-15)
    np.testing.assert_allclose(grouped["scores"][0, 1],
                               vb.bootstrap_rdm_correlation(model, neural2, "spearman", n_bootstrap=30, seed=42)[3], atol=1e-15)


def test_small_and_degenerate_sizes(dev):
    tiny = vb.compute_rdm(torch.randn(3, 5).to(dev))
    assert -1.0 <= vb.compute_rdm_correlation(tiny, tiny.clone(), correlation="Spearman") <= 1.0
    p, lo, hi, s = vb.bootstrap_rdm_correlation(tiny, tiny, "spearman", n_bootstrap=4)     # m = 2 -> one pair -> NaN
    assert len(s) == 4 and all(math.isnan(v) for v in s)
    ten = vb.compute_rdm(torch.randn(10, 6).to(dev))
    p, lo, hi, s = vb.bootstrap_rdm_correlation(ten, ten, "spearman", n_bootstrap=1)
    assert len(s) == 1 and s[0] == pytest.approx(1.0)
    mapped = vb._concept_average_exact(torch.arange(12.0).reshape(6, 2), ["a", "b", "c", "d", "e", "f"],
                                       vb.AlignmentData({}, torch.zeros(2, 1), ["x", "y"], {"x": ["a", "c"], "y": ["f", "zz"]}))
    assert mapped.tolist() == [[2.0, 3.0], [10.0, 11.0]]
