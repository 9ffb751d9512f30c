"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures.

Bars (BASELINE.json north_star): ranks bit-exact given the same RDM; RDM entries within 1e-5
absolute; RSA scores and every bootstrap replicate within 1e-5.
"""
import json
import logging
import math
import os

import numpy as np
import pytest
import torch

import visreps_b200 as vb
from visreps_b200 import rsa as vrsa
from oracle import cases, rsa_oracle as orc

pytestmark = pytest.mark.gpu
RDM_TOL = 1e-5
SCORE_TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


# ------------------------------------------------------------------ RDM
@pytest.mark.parametrize("name", list(cases.RDM_CASES))
def test_rdm_matches_oracle_and_fp64_truth(name, dev, golden_dir):
    x = cases.rdm_input(name)
    keep = x.clone()
    got_cpu_in = vb.compute_rdm(x)                    # CPU tensor in -> CPU tensor out (rsa.py device rule)
    assert got_cpu_in.device.type == "cpu" and got_cpu_in.dtype == torch.float32
    assert torch.equal(x, keep), "input mutated"
    xd = x.to(dev)
    got = vb.compute_rdm(xd)
    assert got.device == xd.device and torch.equal(got.cpu(), got_cpu_in)
    g = got.cpu()
    ref = orc.rdm(x)
    assert float((g - ref).abs().max()) <= RDM_TOL
    assert float(np.abs(g.numpy().astype(np.float64) - orc.rdm_truth_f64(x)).max()) <= RDM_TOL
    assert torch.equal(g, g.T), "RDM must be exactly symmetric"
    assert float(g.diagonal().abs().max()) == 0.0
    assert float(g.min()) >= 0.0 and float(g.max()) <= 2.0
    gold = np.load(os.path.join(golden_dir, "rdm.npz"))
    if f"{name}/pearson" in gold:
        assert float(np.abs(g.numpy() - gold[f"{name}/pearson"]).max()) <= RDM_TOL
    else:
        assert float(np.abs(orc.triu_vector(g)[::97] - gold[f"{name}/pearson_triu_stride97"]).max()) <= RDM_TOL


def test_rdm_known_answers_and_edge_rows(dev):
    simple = torch.tensor([[1., 2., 3.], [4., 5., 6.], [1., 2., 3.], [7., 8., 9.]])
    r = vb.compute_rdm(simple.to(dev)).cpu()
    assert float((r - orc.rdm(simple)).abs().max()) <= 1e-6
    assert abs(float(r[0, 2])) < 1e-4
    anti = torch.tensor([[1., 2., 3., 4.], [4., 3., 2., 1.], [1., 1., 1., 1.], [2., 4., 6., 9.]])
    # constant row: std = sqrt(0 + 1e-12) = 1e-6, which the reference's guard (std < 1e-11, rsa.py:84-87)
    # does NOT catch; its correlations are 0/(..) = 0, so its distances are exactly 1.
    ra = vb.compute_rdm(anti.to(dev)).cpu()
    assert ra[2].tolist() == [1.0, 1.0, 0.0, 1.0]
    assert torch.isfinite(ra).all() and abs(float(ra[0, 1]) - 2.0) < 1e-4
    assert float((ra - orc.rdm(anti)).abs().max()) <= 1e-6
    # n = 1 and n = 0
    one = vb.compute_rdm(torch.randn(1, 7).to(dev))
    assert one.shape == (1, 1) and float(one[0, 0]) == 0.0
    assert vb.compute_rdm(torch.empty(0, 5).to(dev)).shape == (0, 0)
    # other float dtypes are accepted like the reference's .float()
    x = cases.rdm_input("n64_d48")
    r16 = vb.compute_rdm(x.half().to(dev)).cpu()
    assert float((r16 - orc.rdm(x.half())).abs().max()) <= RDM_TOL
    # strided rows (pitch > width) are read in place
    wide = torch.zeros(64, 80, device=dev)
    wide[:, :48] = x.to(dev)
    assert torch.equal(vb.compute_rdm(wide[:, :48]).cpu(), vb.compute_rdm(x.to(dev)).cpu())
    with pytest.raises(ValueError):
        vb.compute_rdm(x.to(dev), correlation="cosine")


def test_rank_rows_and_spearman_rdm(dev):
    x = cases.rdm_input("n64_d48") + torch.from_numpy(np.random.RandomState(5).standard_normal((64, 48)).astype(np.float32))
    assert torch.equal(vb._rank(x.to(dev)).cpu(), orc.ordinal_row_ranks(x))       # no ties: bit-exact
    t = torch.tensor([[3., 1., 2.], [1., 3., 2.]])
    assert torch.equal(vb._rank(t.to(dev)).cpu(), torch.tensor([[2., 0., 1.], [0., 2., 1.]]))
    ties = torch.tensor([[1., 1., 3., 1., -0.0, 0.0]])
    assert vb._rank(ties.to(dev)).cpu().tolist() == [[2., 3., 5., 4., 0., 1.]]     # ties split by position
    big = torch.from_numpy(np.random.RandomState(6).standard_normal((3, 10000)).astype(np.float32))
    assert torch.equal(vb._rank(big.to(dev)).cpu(), orc.ordinal_row_ranks(big))
    rs = vb.compute_rdm(x.to(dev), correlation="Spearman").cpu()
    assert float((rs - orc.rdm(x, correlation="Spearman")).abs().max()) <= RDM_TOL


# ------------------------------------------------------------------ ranks (bit-exact)
def _full_masks(n, dev):
    return vrsa.build_masks(None, n, dev, with_full=True)


@pytest.mark.parametrize("n,levels,seed", [(24, 7, 301), (90, 40, 302), (33, 3, 303), (300, 1000, 304), (65, 1, 305)])
def test_tie_averaged_ranks_are_bit_exact(n, levels, seed, dev):
    N = n * (n - 1) // 2
    va, _ = cases.tied_vectors(N, levels, seed)
    if levels == 1:
        va[:] = 0.5                                  # one run spanning every 32-group
    m = cases.vector_to_rdm(va, n)
    devv, ss, nk = vrsa.subsample_rank(m.to(dev), _full_masks(n, dev))
    got = devv[0].cpu().numpy().astype(np.int64) + (N + 1)
    np.testing.assert_array_equal(got, orc.rank2x(va))
    assert int(nk[0]) == N
    assert float(ss[0]) == float(((got - (N + 1)) ** 2).sum())


@pytest.mark.parametrize("name", ["p64", "p220", "p1000"])
def test_subsample_ranks_are_bit_exact(name, dev):
    x, y = cases.pair_inputs(name)
    n = x.shape[0]
    rdm = vb.compute_rdm(x.to(dev))
    v = orc.triu_vector(rdm)
    idx = vrsa.draw_index_sets(n, 3, 42)
    masks = vrsa.build_masks(idx, n, dev, with_full=True)
    devv, ss, nk = vrsa.subsample_rank(rdm, masks)
    devv = devv.cpu().numpy().astype(np.int64)
    iu = np.triu_indices(n, 1)
    for r in range(4):
        keep = np.ones(n, bool) if r == 0 else np.isin(np.arange(n), idx[r - 1])
        kp = keep[iu[0]] & keep[iu[1]]
        Ns = int(kp.sum())
        assert int(nk[r]) == Ns
        np.testing.assert_array_equal(devv[r][kp] + (Ns + 1), orc.rank2x(v[kp]))
        assert not devv[r][~kp].any()
        assert int(devv[r].sum()) == 0


# ------------------------------------------------------------------ comparisons
@pytest.mark.parametrize("name", list(cases.PAIR_CASES))
def test_compare_matches_oracle_and_golden(name, dev, golden_dir):
    gold = np.load(os.path.join(golden_dir, "compare.npz"))
    x, y = cases.pair_inputs(name)
    a_ref, b_ref = orc.rdm(x), orc.rdm(y)
    for method in ("Spearman", "Pearson"):
        # same RDMs in (reference-built), CPU tensors like the reference's callers pass
        s = vb.compute_rdm_correlation(a_ref, b_ref, correlation=method)
        assert isinstance(s, float)
        assert abs(s - orc.rdm_compare(a_ref, b_ref, correlation=method)) <= SCORE_TOL
        assert abs(s - float(gold[f"{name}/{method.lower()}"])) <= SCORE_TOL
        # end to end: our RDMs -> our comparison, still within the score bar of the reference
        a, b = vb.compute_rdm(x.to(dev)), vb.compute_rdm(y.to(dev))
        assert abs(vb.compute_rdm_correlation(a, b, correlation=method) - float(gold[f"{name}/{method.lower()}"])) <= SCORE_TOL


@pytest.mark.parametrize("tag,n,levels,seed", [("ties24", 24, 7, 301), ("ties90", 90, 40, 302)])
def test_compare_with_heavy_ties(tag, n, levels, seed, dev, golden_dir):
    gold = np.load(os.path.join(golden_dir, "compare.npz"))
    va, vbv = cases.tied_vectors(n * (n - 1) // 2, levels, seed)
    a, b = cases.vector_to_rdm(va, n), cases.vector_to_rdm(vbv, n)
    for method in ("Spearman", "Pearson"):
        s = vb.compute_rdm_correlation(a.to(dev), b.to(dev), correlation=method)
        assert abs(s - float(gold[f"{tag}/{method.lower()}"])) <= SCORE_TOL


@pytest.mark.parametrize("name", ["p64", "p220"])
def test_kendall_tau_a_matches_oracle_and_golden(name, dev, golden_dir):
    gold = np.load(os.path.join(golden_dir, "compare.npz"))
    x, y = cases.pair_inputs(name)
    a, b = orc.rdm(x), orc.rdm(y)
    s = vb.compute_rdm_correlation(a, b, correlation="Kendall")
    assert abs(s - float(gold[f"{name}/kendall"])) <= SCORE_TOL
    assert abs(vb.compute_rdm_correlation(a, b) - s) == 0.0                     # Kendall is the default
    n = a.shape[0]
    idx = orc.draw_subsamples(n, 5, 42)
    point, lo, hi, scores = vb.bootstrap_rdm_correlation(a, b, "kendall", n_bootstrap=5, seed=42)
    np.testing.assert_allclose(scores, orc.bootstrap_compare_with_indices(a, b, "kendall", idx), atol=SCORE_TOL, rtol=0)
    assert abs(point - s) <= 1e-15


def test_kendall_ties_and_hand_counts(dev, golden_dir):
    gold = np.load(os.path.join(golden_dir, "compare.npz"))
    for tag, n, levels, seed in [("ties24", 24, 7, 301), ("ties90", 90, 40, 302)]:
        va, vbv = cases.tied_vectors(n * (n - 1) // 2, levels, seed)
        a, b = cases.vector_to_rdm(va, n), cases.vector_to_rdm(vbv, n)
        assert abs(vb.compute_rdm_correlation(a.to(dev), b.to(dev), correlation="Kendall") - float(gold[f"{tag}/kendall"])) <= SCORE_TOL
    # hand-counted examples on a 3-stimulus RDM (3 pairs): x = (1,2,3)
    mk = lambda v: cases.vector_to_rdm(np.asarray(v, dtype=np.float32), 3)
    assert vb.compute_rdm_correlation(mk([1, 2, 3]), mk([1, 3, 2]), correlation="Kendall") == pytest.approx(1 / 3)
    assert vb.compute_rdm_correlation(mk([1, 2, 3]), mk([3, 1, 2]), correlation="Kendall") == pytest.approx(-1 / 3)
    assert vb.compute_rdm_correlation(mk([1, 2, 3]), mk([3, 2, 1]), correlation="Kendall") == pytest.approx(-1.0)
    # ties lower |tau-a| (no tie correction): x = (1,1,2), y = (1,2,3): C = 2, D = 0, n0 = 3
    assert vb.compute_rdm_correlation(mk([1, 1, 2]), mk([1, 2, 3]), correlation="Kendall") == pytest.approx(2 / 3)


def test_compare_edge_semantics(dev):
    x = torch.randn(20, 10)
    r = vb.compute_rdm(x.to(dev))
    for method in ("Spearman", "Pearson"):
        assert vb.compute_rdm_correlation(r, r, correlation=method) == pytest.approx(1.0, abs=1e-9)
    r2 = 2.0 - r
    r2.fill_diagonal_(0.0)
    assert vb.compute_rdm_correlation(r, r2, correlation="Spearman") == pytest.approx(-1.0, abs=1e-9)
    assert math.isnan(vb.compute_rdm_correlation(torch.zeros(1, 1), torch.zeros(1, 1), correlation="Spearman"))
    assert math.isnan(vb.compute_rdm_correlation(torch.ones(6, 6).to(dev), r[:6, :6], correlation="Spearman"))  # constant
    assert math.isnan(vb.compute_rdm_correlation(torch.ones(6, 6).to(dev), r[:6, :6], correlation="Pearson"))
    two = torch.tensor([[0., 0.3], [0.3, 0.]])
    assert math.isnan(vb.compute_rdm_correlation(two.to(dev), two.to(dev), correlation="Spearman"))       # one pair
    nanm = r.clone()
    nanm[2, 5] = float("nan")
    nanm[5, 2] = float("nan")
    assert math.isnan(vb.compute_rdm_correlation(nanm, r, correlation="Spearman"))                     # NaN propagates
    with pytest.raises(ValueError):
        vb.compute_rdm_correlation(torch.zeros(5, 5), torch.zeros(6, 6))
    with pytest.raises(ValueError):
        vb.compute_rdm_correlation(r, r, correlation="cosine")
    assert vb.compute_rdm_correlation(r, r) == pytest.approx(1.0, abs=1e-12)   # default is Kendall (rsa.py:97)
    assert math.isnan(vb.compute_rdm_correlation(torch.ones(6, 6).to(dev), r[:6, :6], correlation="Kendall"))
    # negative values, -0.0 and wide dynamic range sort correctly
    rs = np.random.RandomState(9)
    n = 40
    N = n * (n - 1) // 2
    va = (rs.standard_normal(N) * 10 ** rs.uniform(-20, 20, N)).astype(np.float32)
    vbv = (va * rs.choice([-1, 1], N) + rs.standard_normal(N)).astype(np.float32)
    va[:5] = [-0.0, 0.0, -1e-38, 1e-38, 0.0]
    a, b = cases.vector_to_rdm(va, n), cases.vector_to_rdm(vbv, n)
    assert abs(vb.compute_rdm_correlation(a, b, correlation="Spearman") - orc.rdm_compare(a, b, correlation="Spearman")) <= SCORE_TOL


# ------------------------------------------------------------------ bootstrap
@pytest.mark.parametrize("name,method", [("p64", "spearman"), ("p64", "pearson"), ("p220", "spearman"),
                                         ("p220", "pearson"), ("p1000", "spearman"), ("p1000", "pearson")])
def test_bootstrap_matches_golden_replicates(name, method, dev, golden_dir):
    gold = np.load(os.path.join(golden_dir, "compare.npz"))
    x, y = cases.pair_inputs(name)
    a_ref, b_ref = orc.rdm(x), orc.rdm(y)
    nb = cases.N_GOLDEN_REPLICATES[name]
    point, lo, hi, scores = vb.bootstrap_rdm_correlation(a_ref, b_ref, method, n_bootstrap=nb, seed=42)
    scores = np.asarray(scores)
    assert scores.shape == (nb,)
    np.testing.assert_allclose(scores, gold[f"{name}/boot_{method}"], atol=SCORE_TOL, rtol=0)
    assert abs(point - float(gold[f"{name}/{method}"])) <= SCORE_TOL
    assert lo == float(np.percentile(scores, 2.5)) and hi == float(np.percentile(scores, 97.5))
    # from our own RDMs too
    a, b = vb.compute_rdm(x.to(dev)), vb.compute_rdm(y.to(dev))
    _, _, _, s2 = vb.bootstrap_rdm_correlation(a, b, method, n_bootstrap=nb, seed=42)
    np.testing.assert_allclose(np.asarray(s2), gold[f"{name}/boot_{method}"], atol=SCORE_TOL, rtol=0)


def test_bootstrap_against_oracle_loop_and_grouping(dev):
    x, y = cases.pair_inputs("p220")
    n = x.shape[0]
    rs = np.random.RandomState(11)
    models = [orc.rdm(x), orc.rdm(x[:, :40]), orc.rdm(torch.from_numpy(rs.standard_normal((n, 30)).astype(np.float32)))]
    neurals = [orc.rdm(y), orc.rdm(y[:, ::2])]
    B = 12
    idx = orc.draw_subsamples(n, B, 42)
    for method in ("spearman", "pearson"):
        res = vb.bootstrap_rdm_correlation_grouped(models, neurals, method, n_bootstrap=B, seed=42)
        assert res["scores"].shape == (3, 2, B) and res["point"].shape == (3, 2)
        for i, m in enumerate(models):
            for j, q in enumerate(neurals):
                want = orc.bootstrap_compare_with_indices(m, q, method, idx)
                np.testing.assert_allclose(res["scores"][i, j], want, atol=SCORE_TOL, rtol=0)
                assert abs(res["point"][i, j] - orc.rdm_compare(m, q, correlation=method.capitalize())) <= SCORE_TOL
        # a replicate slice (multi-GPU shard) reproduces the same entries
        part = vb.bootstrap_rdm_correlation_grouped(models, neurals, method, n_bootstrap=B, seed=42, replicate_range=(5, 9))
        np.testing.assert_array_equal(part["scores"], res["scores"][:, :, 5:9])
    # n_bootstrap = 1 works (reference tests :2058-2069)
    p, lo, hi, s = vb.bootstrap_rdm_correlation(models[0], neurals[0], "spearman", n_bootstrap=1)
    assert len(s) == 1 and lo == hi == s[0]


def test_large_n_global_table_path(dev):
    """n = 1500 (N = 1,124,250 pairs): the count tables no longer fit shared memory and the sums leave the
    exact-int64 regime, so the global-table / fp64-accumulator variants of the kernels run."""
    rs = np.random.RandomState(31)
    n = 1500
    z = rs.standard_normal((n, 24)).astype(np.float32)
    x = torch.from_numpy(np.maximum(z @ rs.standard_normal((24, 200)).astype(np.float32) + rs.standard_normal((n, 200)).astype(np.float32), 0))
    y = torch.from_numpy((z @ rs.standard_normal((24, 90)) + 2 * rs.standard_normal((n, 90))).astype(np.float32))
    a, b = vb.compute_rdm(x.to(dev)), vb.compute_rdm(y.to(dev))
    assert float((a.cpu() - orc.rdm(x)).abs().max()) <= RDM_TOL
    idx = vrsa.draw_index_sets(n, 2, 42)
    res = vb.bootstrap_rdm_correlation_grouped([a], [b], "spearman", index_sets=idx)
    want = orc.bootstrap_compare_with_indices(a.cpu(), b.cpu(), "spearman", idx)
    np.testing.assert_allclose(res["scores"][0, 0], want, atol=SCORE_TOL, rtol=0)
    assert abs(res["point"][0, 0] - orc.rdm_compare(a.cpu(), b.cpu(), correlation="Spearman")) <= SCORE_TOL
    # ranks stay bit-exact on this path too
    masks = vrsa.build_masks(idx[:1], n, dev, with_full=True)
    devv, ss, nk = vrsa.subsample_rank(a, masks)
    v = orc.triu_vector(a)
    iu = np.triu_indices(n, 1)
    keep = np.isin(np.arange(n), idx[0])
    kp = keep[iu[0]] & keep[iu[1]]
    got = devv[1].cpu().numpy().astype(np.int64)
    np.testing.assert_array_equal(got[kp] + (int(kp.sum()) + 1), orc.rank2x(v[kp]))
    pe = vb.compute_rdm_correlation(a, b, correlation="Pearson")
    assert abs(pe - orc.rdm_compare(a.cpu(), b.cpu(), correlation="Pearson")) <= SCORE_TOL


def test_things_shape_behaviour_rdm(dev):
    """BASELINE config 3 shape: 1854 concepts, 66-d behavioural embedding (non-negative) vs model layers.
    d = 66 is not a multiple of the TMA box; N = 1,717,731 pairs uses the global-table, fp64-accumulator kernels."""
    rs = np.random.RandomState(41)
    n = 1854
    z = rs.standard_normal((n, 20)).astype(np.float32)
    beh = torch.from_numpy(np.maximum(z @ rs.standard_normal((20, 66)).astype(np.float32), 0))
    layers = [torch.from_numpy(np.maximum(z @ rs.standard_normal((20, d)).astype(np.float32) + rs.standard_normal((n, d)).astype(np.float32), 0))
              for d in (96, 513)]
    rb = vb.compute_rdm(beh.to(dev))
    assert float((rb.cpu() - orc.rdm(beh)).abs().max()) <= RDM_TOL
    idx = vrsa.draw_index_sets(n, 2, 42)
    res = vb.rsa_grouped(layers, [beh], "spearman", n_bootstrap=2, seed=42)
    assert res["scores"].shape == (2, 1, 2)
    for i, x in enumerate(layers):
        ra, rbb = orc.rdm(x), orc.rdm(beh)
        assert abs(res["point"][i, 0] - orc.rdm_compare(ra, rbb, correlation="Spearman")) <= SCORE_TOL
        if i == 0:
            want = orc.bootstrap_compare_with_indices(ra, rbb, "spearman", idx[:1])
            assert abs(res["scores"][0, 0, 0] - want[0]) <= SCORE_TOL


def test_replicate_chunking_with_a_small_workspace(dev):
    """vr_bootstrap_spearman processes the replicates in chunks when the caller's workspace is small:
    same scores as one pass."""
    import ctypes as C
    from visreps_b200 import _lib
    lib = _lib.load()
    x, y = cases.pair_inputs("p64")
    n, R = 64, 150
    a, b = vrsa.PresortedRDM(vb.compute_rdm(x.to(dev))), vrsa.PresortedRDM(vb.compute_rdm(y.to(dev)))
    masks = vrsa.build_masks(vrsa.draw_index_sets(n, R - 1, 42), n, dev, with_full=True)
    full = vrsa.compare_grouped([a], [b], "spearman", masks)
    small_bytes = lib.vr_bootstrap_spearman_workspace_bytes(n, 1, 1, 64)       # room for 64 of the 150 replicates
    assert small_bytes < lib.vr_bootstrap_spearman_workspace_bytes(n, 1, 1, R)
    ws = torch.empty(small_bytes, dtype=torch.uint8, device=dev)
    out = torch.full((1, 1, R), -7.0, dtype=torch.float64, device=dev)
    pa, pb = (C.c_void_p * 1)(a.blob.data_ptr()), (C.c_void_p * 1)(b.blob.data_ptr())
    _lib.check(lib.vr_bootstrap_spearman(pa, 1, pb, 1, n, masks.data_ptr(), R, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                         _lib.stream_ptr()), "vr_bootstrap_spearman")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), full.cpu().numpy())
    # a workspace below the minimum is refused with a message, nothing is launched
    rc = lib.vr_bootstrap_spearman(pa, 1, pb, 1, n, masks.data_ptr(), R, out.data_ptr(), ws.data_ptr(), 1024, _lib.stream_ptr())
    assert rc == _lib.VR_ERR_WORKSPACE and b"workspace" in lib.vr_last_error_string()


def test_very_large_n_paths_agree(dev):
    """n = 3700 (N = 6.8 M pairs, global-table kernels, fp64 accumulators): a replicate must equal the full
    comparison of the gathered sub-RDMs (N = 5.5 M), and self-comparison must be 1 (no CPU loop at this size)."""
    g = torch.Generator(device=dev).manual_seed(3)
    n = 3700
    z = torch.randn(n, 32, device=dev, generator=g)
    a = vb.compute_rdm(torch.relu(z @ torch.randn(32, 256, device=dev, generator=g) + torch.randn(n, 256, device=dev, generator=g)))
    b = vb.compute_rdm(z @ torch.randn(32, 100, device=dev, generator=g) + 2 * torch.randn(n, 100, device=dev, generator=g))
    idx = vrsa.draw_index_sets(n, 2, 42)
    res = vb.bootstrap_rdm_correlation_grouped([a], [b], "spearman", index_sets=idx)
    for r in range(2):
        t = torch.from_numpy(idx[r].astype(np.int64)).to(dev)
        direct = vb.compute_rdm_correlation(a[t][:, t], b[t][:, t], correlation="Spearman")
        assert abs(direct - res["scores"][0, 0, r]) <= 1e-12
    self_s = vb.bootstrap_rdm_correlation_grouped([a], [a], "spearman", index_sets=idx)
    np.testing.assert_allclose(self_s["scores"], 1.0, atol=1e-12)
    assert 0.0 < res["point"][0, 0] < 1.0


def test_grouped_with_more_than_one_materialised_group(dev):
    """10 model x 9 neural RDMs: the smaller side is materialised in groups of 8, so two groups."""
    rs = np.random.RandomState(21)
    n, B = 64, 5
    base = rs.standard_normal((n, 12)).astype(np.float32)
    mk = lambda k: orc.rdm(torch.from_numpy((base @ rs.standard_normal((12, 20 + k)) + rs.standard_normal((n, 20 + k))).astype(np.float32)))
    models = [mk(k) for k in range(10)]
    neurals = [mk(k + 3) for k in range(9)]
    idx = orc.draw_subsamples(n, B, 42)
    res = vb.bootstrap_rdm_correlation_grouped(models, neurals, "spearman", n_bootstrap=B, seed=42)
    for i in (0, 7, 8, 9):
        for j in (0, 4, 7, 8):
            want = orc.bootstrap_compare_with_indices(models[i], neurals[j], "spearman", idx)
            np.testing.assert_allclose(res["scores"][i, j], want, atol=SCORE_TOL, rtol=0)
            assert abs(res["point"][i, j] - orc.rdm_compare(models[i], neurals[j], correlation="Spearman")) <= SCORE_TOL
    # transposed roles (fewer models than neural RDMs) give the transposed result
    res_t = vb.bootstrap_rdm_correlation_grouped(neurals, models, "spearman", n_bootstrap=B, seed=42)
    np.testing.assert_allclose(res_t["scores"], np.transpose(res["scores"], (1, 0, 2)), atol=1e-15, rtol=0)


def test_rsa_grouped_pipeline_host_and_device_inputs(dev):
    """rsa_grouped = the scoring block of _eval_rsa for all (layer, ROI) pairs; host inputs take the
    copy/compute-overlapped path, device inputs the batched one: identical results, both equal to the oracle loop."""
    rs = np.random.RandomState(17)
    n, B = 96, 9
    z = rs.standard_normal((n, 10)).astype(np.float32)
    feats = [torch.from_numpy(np.maximum(z @ rs.standard_normal((10, d)).astype(np.float32) + rs.standard_normal((n, d)).astype(np.float32), 0))
             for d in (300, 40, 1200)]
    feats[1] = feats[1].reshape(n, 2, 4, 5)                      # >2-D activations are flattened (evals.py:321)
    neur = [torch.from_numpy((z @ rs.standard_normal((10, v)) + rs.standard_normal((n, v))).astype(np.float32)) for v in (25, 60, 33, 80)]
    host = vb.rsa_grouped([f.pin_memory() for f in feats], neur, "spearman", n_bootstrap=B, seed=42)
    devr = vb.rsa_grouped([f.to(dev) for f in feats], [y.to(dev) for y in neur], "spearman", n_bootstrap=B, seed=42)
    for key in ("point", "scores", "ci_low", "ci_high"):
        np.testing.assert_array_equal(host[key], devr[key])
    assert host["scores"].shape == (3, 4, B)
    idx = orc.draw_subsamples(n, B, 42)
    for i, f in enumerate(feats):
        for j, y in enumerate(neur):
            ra, rb = orc.rdm(f.flatten(start_dim=1)), orc.rdm(y)
            np.testing.assert_allclose(host["scores"][i, j], orc.bootstrap_compare_with_indices(ra, rb, "spearman", idx), atol=SCORE_TOL, rtol=0)
            assert abs(host["point"][i, j] - orc.rdm_compare(ra, rb, correlation="Spearman")) <= SCORE_TOL
    np.testing.assert_allclose(host["ci_low"], np.percentile(host["scores"], 2.5, axis=2))
    # more neural than model matrices -> the model side is materialised; and the Pearson route
    swapped = vb.rsa_grouped(neur, feats, "spearman", n_bootstrap=B, seed=42)
    np.testing.assert_allclose(swapped["scores"], np.transpose(host["scores"], (1, 0, 2)), atol=1e-15, rtol=0)
    pear = vb.rsa_grouped(feats, neur, "pearson", n_bootstrap=3, seed=42)
    ra, rb = orc.rdm(feats[2]), orc.rdm(neur[1])
    np.testing.assert_allclose(pear["scores"][2, 1], orc.bootstrap_compare_with_indices(ra, rb, "pearson", idx[:3]), atol=SCORE_TOL, rtol=0)
    nb = vb.rsa_grouped(feats, neur, "spearman", bootstrap=False)
    assert nb["scores"].shape == (3, 4, 0) and "ci_low" not in nb
    np.testing.assert_array_equal(nb["point"], host["point"])


def test_compute_rsa_matches_reference_fixture(dev, golden_dir):
    with open(os.path.join(golden_dir, "compute_rsa.json")) as f:
        gold = json.load(f)["results"]
    tr, te = cases.alignment_case()
    for tag, kw in {"all_train": dict(n_select=None), "subsample": dict(n_select=120)}.items():
        want = gold[f"{tag}/spearman"]
        got = vb.compute_rsa({"compare_method": "spearman"}, vb.AlignmentData(**tr), vb.AlignmentData(**te),
                             bootstrap=True, n_bootstrap=25, seed=7, **kw)[0]
        assert list(got.keys()) == list(want.keys())
        assert got["layer"] == want["layer"] == "good" and got["analysis"] == "rsa" and got["compare_method"] == "spearman"
        assert abs(got["score"] - want["score"]) <= SCORE_TOL
        np.testing.assert_allclose(got["bootstrap_scores"], want["bootstrap_scores"], atol=SCORE_TOL, rtol=0)
        assert abs(got["ci_low"] - want["ci_low"]) <= SCORE_TOL and abs(got["ci_high"] - want["ci_high"]) <= SCORE_TOL
        for g, w in zip(got["layer_selection_scores"], want["layer_selection_scores"]):
            assert g["layer"] == w["layer"] and abs(g["score"] - w["score"]) <= SCORE_TOL
    # no bootstrap -> CI fields None, no bootstrap_scores key (rsa.py:229-279)
    nb = vb.compute_rsa({"compare_method": "spearman"}, vb.AlignmentData(**tr), vb.AlignmentData(**te), bootstrap=False)[0]
    assert nb["ci_low"] is None and nb["ci_high"] is None and "bootstrap_scores" not in nb
    # re_extract_fn is used for the evaluation RDM of the chosen layer (rsa.py:212-215)
    calls = []

    def re_extract(layer, ids):
        calls.append((layer, ids))
        return te["activations"][layer], ids
    r2 = vb.compute_rsa({"compare_method": "spearman"}, vb.AlignmentData(**tr), vb.AlignmentData(**te), bootstrap=False,
                        re_extract_fn=re_extract)[0]
    assert calls == [("good", te["stimulus_ids"])] and abs(r2["score"] - nb["score"]) <= 1e-12
    # compare_method = kendall (config-valid, utils.py:518): tau-a for selection, point estimate and replicates
    kw = dict(bootstrap=True, n_bootstrap=4, seed=3)
    got = vb.compute_rsa({"compare_method": "kendall"}, vb.AlignmentData(**tr), vb.AlignmentData(**te), **kw)[0]
    want = orc.rsa_traintest({"compare_method": "kendall"}, orc.Split(**tr), orc.Split(**te), **kw)[0]
    assert got["layer"] == want["layer"] and abs(got["score"] - want["score"]) <= SCORE_TOL
    np.testing.assert_allclose(got["bootstrap_scores"], want["bootstrap_scores"], atol=SCORE_TOL, rtol=0)


# ------------------------------------------------------------------ full-size properties (no oracle loop)
def test_nsd_shape_properties(dev):
    """BASELINE config 2/3 sizes: checks that do not need the 0.25 s/replicate CPU path."""
    x, y = cases.pair_inputs("p1000")
    n = 1000
    a, b = vb.compute_rdm(x.to(dev)), vb.compute_rdm(y.to(dev))
    B = 40
    idx = vrsa.draw_index_sets(n, B, 42)
    res = vb.bootstrap_rdm_correlation_grouped([a], [b], "spearman", index_sets=idx)
    s = res["scores"][0, 0]
    # (1) sub-index == direct: a replicate equals the full comparison of the gathered sub-RDMs (reference tests :1390-1412)
    for r in (0, 17, 39):
        t = torch.from_numpy(idx[r].astype(np.int64)).to(dev)
        direct = vb.compute_rdm_correlation(a[t][:, t], b[t][:, t], correlation="Spearman")
        assert abs(direct - s[r]) <= 1e-12
    # (2) order inside a draw is irrelevant; (3) determinism
    shuffled = np.stack([np.random.RandomState(r).permutation(row) for r, row in enumerate(idx)])
    again = vb.bootstrap_rdm_correlation_grouped([a], [b], "spearman", index_sets=shuffled)["scores"][0, 0]
    np.testing.assert_array_equal(s, again)
    # (4) self-comparison is exactly 1 for every replicate; symmetry in the arguments
    self_s = vb.bootstrap_rdm_correlation_grouped([a], [a], "spearman", index_sets=idx)
    np.testing.assert_allclose(self_s["scores"], 1.0, atol=1e-12)
    swapped = vb.bootstrap_rdm_correlation_grouped([b], [a], "spearman", index_sets=idx)["scores"][0, 0]
    np.testing.assert_allclose(swapped, s, atol=1e-15)
    # (5) scores are plausible and spread around the point estimate
    assert abs(np.median(s) - res["point"][0, 0]) < 0.02 and s.std() > 0


def test_tied_and_tie_free_vectors_in_one_launch(dev):
    """Vectors without ties take the 2 C(s) + 1 sweep, vectors with ties the general one; which is which is decided
    on the device from the presort header, so one grouped call may hold both kinds on either side."""
    n, B = 130, 9
    N = n * (n - 1) // 2
    rs = np.random.RandomState(77)

    def tied(levels, seed):
        return cases.vector_to_rdm(cases.tied_vectors(N, levels, seed)[0], n)

    def smooth(seed):
        return orc.rdm(torch.from_numpy(np.random.RandomState(seed).standard_normal((n, 50)).astype(np.float32)))

    one_tie = smooth(5).clone()
    one_tie[3, 7] = one_tie[7, 3] = one_tie[90, 120]; one_tie[120, 90] = one_tie[90, 120]      # a single tied pair
    models = [smooth(1), tied(5, 2), smooth(3), one_tie, tied(2000, 4)]
    neurals = [tied(9, 6), smooth(7), smooth(8)]
    idx = orc.draw_subsamples(n, B, 42)
    res = vb.bootstrap_rdm_correlation_grouped(models, neurals, "spearman", n_bootstrap=B, seed=42)
    for i, m in enumerate(models):
        for j, q in enumerate(neurals):
            want = orc.bootstrap_compare_with_indices(m, q, "spearman", idx)
            np.testing.assert_allclose(res["scores"][i, j], want, atol=SCORE_TOL, rtol=0)
            assert abs(res["point"][i, j] - orc.rdm_compare(m, q, correlation="Spearman")) <= SCORE_TOL
    # ranks and the sum of squares of a tie-free vector (closed form on the device) are bit-exact
    masks = vrsa.build_masks(idx, n, dev, with_full=True)
    for m in (models[0], models[3]):
        devv, ss, nk = vrsa.subsample_rank(m.to(dev), masks)
        devv = devv.cpu().numpy().astype(np.int64)
        v = orc.triu_vector(m)
        iu = np.triu_indices(n, 1)
        for r in range(B + 1):
            keep = np.ones(n, bool) if r == 0 else np.isin(np.arange(n), idx[r - 1])
            kp = keep[iu[0]] & keep[iu[1]]
            Ns = int(kp.sum())
            np.testing.assert_array_equal(devv[r][kp] + (Ns + 1), orc.rank2x(v[kp]))
            assert not devv[r][~kp].any()
            assert float(ss[r]) == float((devv[r] ** 2).sum())


def test_eval_blocks_match_the_reference_loops(dev):
    """visreps_b200.evals restates the (region, subject) loops of visreps/evals.py:249-287 and :333-398 as grouped
    device jobs; the oracle runs the loops as written."""
    from visreps_b200 import evals as vevals
    rs = np.random.RandomState(2024)
    g = lambda *s: torch.from_numpy(rs.standard_normal(s).astype(np.float32))
    n_train, n_test, n_select, B = 150, 60, 100, 25
    layers = ["conv1", "conv2", "fc1"]
    pairs = [("V1", 0), ("V1", 1), ("ventral", 0), ("ventral", 1)]
    latent_tr, latent_te = g(n_train, 12), g(n_test, 12)
    mix = {l: g(12, d) for l, d in zip(layers, (300, 80, 40))}
    noise = {"conv1": 3.0, "conv2": 0.3, "fc1": 1.0}
    train_acts = {k: {l: (latent_tr @ mix[l] + noise[l] * g(n_train, mix[l].shape[1])).reshape(n_train, -1, 4) for l in layers} for k in pairs}
    test_feats = {l: (latent_te @ mix[l] + noise[l] * g(n_test, mix[l].shape[1])).reshape(n_test, 4, -1) for l in layers}
    proj = {k: g(12, 30 + 7 * i) for i, k in enumerate(pairs)}
    train_neural = {k: latent_tr @ proj[k] + 0.5 * g(n_train, proj[k].shape[1]) for k in pairs}
    test_neural = {k: latent_te @ proj[k] + 0.5 * g(n_test, proj[k].shape[1]) for k in pairs}

    got = vevals.eval_rsa_scores(train_acts, train_neural, test_feats, test_neural, "spearman", n_select=n_select,
                                 bootstrap=True, n_bootstrap=B)
    assert [r["layer"] for r in got] and len(got) == len(pairs)
    model_rdms = {l: orc.rdm(test_feats[l].flatten(start_dim=1)) for l in layers}
    for k, r in zip(pairs, got):
        # phase 1 as written: RandomState(42) per pair, first strict maximum wins
        sel_idx = np.random.RandomState(42).choice(n_train, size=n_select, replace=False)
        nrdm = orc.rdm(train_neural[k][sel_idx])
        best, best_score, want_sel = None, -float("inf"), []
        for l in layers:
            s = orc.rdm_compare(orc.rdm(train_acts[k][l][sel_idx].flatten(start_dim=1)), nrdm, correlation="Spearman")
            want_sel.append(s)
            if s > best_score:
                best, best_score = l, s
        assert r["layer"] == best
        assert [d["layer"] for d in r["layer_selection_scores"]] == layers
        np.testing.assert_allclose([d["score"] for d in r["layer_selection_scores"]], want_sel, atol=SCORE_TOL, rtol=0)
        # scoring block as written: RandomState(42) re-created per pair
        neural_rdm = orc.rdm(test_neural[k])
        assert abs(r["score"] - orc.rdm_compare(model_rdms[best], neural_rdm, correlation="Spearman")) <= SCORE_TOL
        want = orc.bootstrap_compare_with_indices(model_rdms[best], neural_rdm, "spearman", orc.draw_subsamples(n_test, B, 42))
        np.testing.assert_allclose(r["bootstrap_scores"], want, atol=SCORE_TOL, rtol=0)
        assert r["ci_low"] == float(np.percentile(r["bootstrap_scores"], 2.5))
        assert r["ci_high"] == float(np.percentile(r["bootstrap_scores"], 97.5))
        assert set(r) == {"layer", "compare_method", "score", "ci_low", "ci_high", "analysis", "layer_selection_scores", "bootstrap_scores"}
        assert isinstance(r["score"], float) and isinstance(r["bootstrap_scores"], list)
    # without bootstrap: CI fields are None and there is no score list (evals.py:350-352, :388-389)
    plain = vevals.score_pairs(test_feats, test_neural, {k: "fc1" for k in pairs}, "pearson")
    assert all(r["ci_low"] is None and r["ci_high"] is None and "bootstrap_scores" not in r and r["layer_selection_scores"] == [] for r in plain)
    for k, r in zip(pairs, plain):
        assert abs(r["score"] - orc.rdm_compare(model_rdms["fc1"], orc.rdm(test_neural[k]), correlation="Pearson")) <= SCORE_TOL
    # n_select >= n_train keeps every stimulus in order (evals.py:261-262)
    assert np.array_equal(vevals.selection_indices(80, 1000), np.arange(80)) and np.array_equal(vevals.selection_indices(80, None), np.arange(80))


def test_kendall_batches_ties_and_small_workspace(dev):
    """More replicates than one batch (64), heavy ties on one side, and the minimum workspace (one replicate per
    batch) must all give the oracle's tau-a."""
    import ctypes as C
    from visreps_b200 import _lib
    n, B = 150, 70
    N = n * (n - 1) // 2
    rs = np.random.RandomState(99)
    a = orc.rdm(torch.from_numpy(rs.standard_normal((n, 40)).astype(np.float32)))
    b = cases.vector_to_rdm(cases.tied_vectors(N, 11, 5)[0], n)
    c = orc.rdm(torch.from_numpy(rs.standard_normal((n, 25)).astype(np.float32)))
    idx = orc.draw_subsamples(n, B, 42)
    for x, y in ((a, b), (a, c), (b, b)):
        p, lo, hi, scores = vb.bootstrap_rdm_correlation(x, y, "kendall", n_bootstrap=B, seed=42)
        want = orc.bootstrap_compare_with_indices(x, y, "kendall", idx)
        np.testing.assert_allclose(scores, want, atol=SCORE_TOL, rtol=0)
        assert abs(p - orc.rdm_compare(x, y, correlation="Kendall")) <= SCORE_TOL
    # the C entry point with its minimum workspace: same numbers, one replicate per batch
    lib = _lib.load()
    masks = vrsa.build_masks(idx, n, dev, with_full=True)
    R = int(masks.shape[0])
    xa, yb = a.to(dev).contiguous(), b.to(dev).contiguous()
    small = torch.empty(lib.vr_kendall_workspace_bytes(n), dtype=torch.uint8, device=dev)
    big = torch.empty(lib.vr_kendall_batch_workspace_bytes(n, R), dtype=torch.uint8, device=dev)
    outs = []
    for ws in (small, big):
        out = torch.empty(R, dtype=torch.float64, device=dev)
        _lib.check(lib.vr_kendall_tau_a(xa.data_ptr(), yb.data_ptr(), n, n, masks.data_ptr(), R, out.data_ptr(),
                                        ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "vr_kendall_tau_a")
        outs.append(out.cpu().numpy())
    np.testing.assert_array_equal(outs[0], outs[1])
    np.testing.assert_allclose(outs[0][1:], orc.bootstrap_compare_with_indices(a, b, "kendall", idx), atol=SCORE_TOL, rtol=0)
    assert lib.vr_kendall_tau_a(xa.data_ptr(), yb.data_ptr(), n, n, masks.data_ptr(), R, out.data_ptr(),
                                small.data_ptr(), small.numel() - 4096, _lib.stream_ptr()) == 2      # VR_ERR_WORKSPACE


def test_pearson_inclusion_exclusion_paths_and_nan_semantics(dev):
    """The Pearson bootstrap sums over the DROPPED stimuli (inclusion-exclusion) and falls back to the kept pairs
    when a mask drops more than 2048 stimuli; a NaN entry makes exactly those replicates NaN that keep its pair."""
    rs = np.random.RandomState(31)
    n = 90
    a = orc.rdm(torch.from_numpy(rs.standard_normal((n, 30)).astype(np.float32)))
    b = orc.rdm(torch.from_numpy(rs.standard_normal((n, 12)).astype(np.float32)))
    c = orc.rdm(torch.from_numpy(rs.standard_normal((n, 7)).astype(np.float32)))
    # arbitrary masks: keep 2 .. n stimuli
    sets = [np.sort(rs.choice(n, size=k, replace=False)) for k in (2, 3, 17, 45, 80, 89, 90)]
    W = (n + 31) // 32
    words = np.zeros((len(sets), W), dtype=np.uint32)
    for r, ks in enumerate(sets):
        for i in ks:
            words[r, i >> 5] |= np.uint32(1) << np.uint32(i & 31)
    masks = torch.from_numpy(words.view(np.int32)).to(dev)
    out = vrsa.compare_grouped([vrsa._RawRDM(a.to(dev)), vrsa._RawRDM(c.to(dev))], [vrsa._RawRDM(b.to(dev))], "pearson", masks).cpu().numpy()
    for r, ks in enumerate(sets):
        for k, m in enumerate((a, c)):
            sub = torch.from_numpy(ks)
            want = orc.rdm_compare(m[sub][:, sub], b[sub][:, sub], correlation="Pearson")
            got = out[k, 0, r]
            assert (math.isnan(want) and math.isnan(got)) or abs(got - want) <= SCORE_TOL, (r, k, got, want)
    # NaN entries: stimulus 5's row/column is NaN in `a`
    bad = a.clone()
    bad[5, :] = float("nan"); bad[:, 5] = float("nan"); bad[5, 5] = 0.0
    keep_all = np.arange(n)
    without5 = np.array([i for i in range(n) if i not in (5, 40)])
    words = np.zeros((2, W), dtype=np.uint32)
    for r, ks in enumerate((keep_all, without5)):
        for i in ks:
            words[r, i >> 5] |= np.uint32(1) << np.uint32(i & 31)
    out = vrsa.compare_grouped([vrsa._RawRDM(bad.to(dev))], [vrsa._RawRDM(b.to(dev))], "pearson",
                               torch.from_numpy(words.view(np.int32)).to(dev)).cpu().numpy()[0, 0]
    assert math.isnan(out[0])
    sub = torch.from_numpy(without5)
    assert abs(out[1] - orc.rdm_compare(a[sub][:, sub], b[sub][:, sub], correlation="Pearson")) <= SCORE_TOL
    # more than 2048 dropped stimuli: the direct sum over the kept pairs
    n2 = 2200
    x = torch.from_numpy(rs.standard_normal((n2, 16)).astype(np.float32))
    y = x + torch.from_numpy(rs.standard_normal((n2, 16)).astype(np.float32))
    ra, rb = vb.compute_rdm(x.to(dev)), vb.compute_rdm(y.to(dev))
    ks = np.sort(rs.choice(n2, size=120, replace=False))
    words = np.zeros((2, (n2 + 31) // 32), dtype=np.uint32)
    for i in ks:
        words[0, i >> 5] |= np.uint32(1) << np.uint32(i & 31)
    words[1, :] = 0xFFFFFFFF
    words[1, -1] = (1 << (n2 % 32)) - 1 if n2 % 32 else 0xFFFFFFFF
    out = vrsa.compare_grouped([vrsa._RawRDM(ra)], [vrsa._RawRDM(rb)], "pearson", torch.from_numpy(words.view(np.int32)).to(dev)).cpu().numpy()[0, 0]
    sub = torch.from_numpy(ks).to(dev)
    iu = np.triu_indices(len(ks), 1)
    va, vbv = ra[sub][:, sub].cpu().numpy()[iu].astype(np.float64), rb[sub][:, sub].cpu().numpy()[iu].astype(np.float64)
    assert abs(out[0] - np.corrcoef(va, vbv)[0, 1]) <= SCORE_TOL
    iu = np.triu_indices(n2, 1)
    assert abs(out[1] - np.corrcoef(ra.cpu().numpy()[iu].astype(np.float64), rb.cpu().numpy()[iu].astype(np.float64))[0, 1]) <= SCORE_TOL


def test_rdm_wide_rows_single_pass_centring(dev):
    """Rows of >= 16384 features are centred in one pass around a sampled shift, with the remainder of the mean
    removed in the epilogue; offsets, block structure and sparsity must not cost accuracy (1e-5 vs fp64)."""
    rs = np.random.RandomState(123)
    n, d = 96, 20000
    base = rs.standard_normal((n, d)).astype(np.float32)
    x = base.copy()
    x[0:16] += 50.0                                            # a large common offset
    x[16:32, : d // 2] = 0.0                                   # block structure: the sample mean is not the mean
    x[16:32, d // 2:] = x[16:32, d // 2:] * 3 + 40.0
    x[32:48] = np.maximum(x[32:48], 0)                         # post-ReLU sparsity
    x[48:56, ::7] += 1000.0                                    # periodic spikes that alias with a strided sample
    x[56] = 3.25                                               # a constant row
    t = torch.from_numpy(x)
    got = vb.compute_rdm(t.to(dev)).cpu().double().numpy()
    xd = x.astype(np.float64)
    xc = xd - xd.mean(1, keepdims=True)
    sd = np.sqrt((xc ** 2).mean(1) + 1e-12)
    truth = 1 - np.clip((xc @ xc.T) / d / (sd[:, None] * sd[None, :] + 1e-12), -1, 1)
    np.fill_diagonal(truth, 0.0)
    assert np.abs(got - truth).max() <= RDM_TOL
    assert np.abs(got - orc.rdm(t).double().numpy()).max() <= RDM_TOL
    assert np.array_equal(got, got.T) and not np.diag(got).any()


def test_encoding_score_bootstrap_matches_oracle(dev):
    """the bootstrap over cached predictions (encoding_score.py:228-240) on the device."""
    from oracle import encoding_oracle as enc
    from visreps_b200 import encoding as venc
    rs = np.random.RandomState(8)
    n, V, B = 220, 300, 40
    y = rs.standard_normal((n, V)).astype(np.float32) * rs.uniform(0.5, 4.0, V).astype(np.float32) + rs.uniform(-3, 3, V).astype(np.float32)
    p = (0.5 * y + rs.standard_normal((n, V)) * 2.0).astype(np.float32)
    rng_o, rng_g = np.random.RandomState(42), np.random.RandomState(42)
    rng_o.permutation(1000); rng_g.permutation(1000)           # the train split drawn first (encoding_score.py:131)
    want = enc.bootstrap_encoding(y, p, n_bootstrap=B, rng=rng_o)
    got = venc.bootstrap_encoding_score(torch.from_numpy(y), torch.from_numpy(p), n_bootstrap=B, rng=rng_g)
    assert abs(got[0] - want[0]) <= SCORE_TOL
    np.testing.assert_allclose(got[3], want[3], atol=SCORE_TOL, rtol=0)
    assert abs(got[1] - want[1]) <= SCORE_TOL and abs(got[2] - want[2]) <= SCORE_TOL
    assert rng_o.randint(1 << 30) == rng_g.randint(1 << 30)    # generators left in step
    assert isinstance(got[3], list) and len(got[3]) == B
    cols = venc.correlation_score(torch.from_numpy(y).to(dev), torch.from_numpy(p).to(dev)).cpu().numpy()
    np.testing.assert_allclose(cols, enc.correlation_score(y, p), atol=SCORE_TOL, rtol=0)
    # a column without variance among the kept rows makes the replicate NaN, like the division by a zero std
    y2 = y.copy(); y2[:, 3] = 1.5
    point, lo, hi, scores = venc.bootstrap_encoding_score(torch.from_numpy(y2), torch.from_numpy(p), n_bootstrap=3)
    assert math.isnan(point) and all(math.isnan(s) for s in scores)
    # strided inputs and many dropped rows (direct sum over the kept rows)
    wide = torch.from_numpy(np.concatenate([y, p], axis=1)).to(dev)
    a, b = wide[:, :V], wide[:, V:]
    np.testing.assert_allclose(venc.correlation_score(a, b).cpu().numpy(), cols, atol=1e-7, rtol=0)
