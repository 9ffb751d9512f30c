"""Host-side logic that needs no GPU: RNG stream fidelity, masks, sharding arithmetic,
and a numpy walk-through of the device ranking algorithm against the oracle."""
import numpy as np
import pytest

from oracle import rsa_oracle as orc
from visreps_b200 import rsa as vrsa


def test_index_sets_follow_the_reference_rng_stream():
    for n, b, seed in [(50, 7, 42), (220, 5, 42), (1000, 3, 7)]:
        mine = vrsa.draw_index_sets(n, b, seed)
        np.testing.assert_array_equal(mine, orc.draw_subsamples(n, b, seed))
        assert mine.shape == (b, int(n * 0.9))
        assert all(len(set(row.tolist())) == mine.shape[1] for row in mine)   # without replacement
    # a shared RandomState keeps being consumed (compute_rsa: selection draw first, rsa.py:176)
    rng = np.random.RandomState(3)
    first = rng.choice(200, size=120, replace=False)
    rest = vrsa.draw_index_sets(50, 4, rng=rng)
    rng2 = np.random.RandomState(3)
    np.testing.assert_array_equal(first, rng2.choice(200, size=120, replace=False))
    np.testing.assert_array_equal(rest, orc.draw_subsamples(50, 4, rng=rng2))


def test_full_mask_bits():
    for n in (1, 31, 32, 33, 1000, 1854):
        m = vrsa.full_mask(n)
        assert m.dtype == np.uint32 and m.size == (n + 31) // 32
        assert sum(bin(int(w)).count("1") for w in m) == n


def _device_algorithm_numpy(v, n, mask):
    """Numpy restatement of presort + rank_kernel (rank.cu): C(s) + C(e) - N_r at the
    pair's canonical position, with the 32-group head tables."""
    N = n * (n - 1) // 2
    iu = np.triu_indices(n, 1)
    key = v.copy()
    key[key == 0] = 0.0                               # -0.0 folds onto +0.0
    order = np.argsort(key, kind="stable")
    sv, si, sj = key[order], iu[0][order], iu[1][order]
    head = np.r_[True, sv[1:] != sv[:-1]]
    hp = np.flatnonzero(head)
    kept = mask[si] & mask[sj]
    C = np.r_[0, np.cumsum(kept)]
    Ns = int(C[-1])
    start = hp[np.searchsorted(hp, np.arange(N), side="right") - 1]
    nxt = np.searchsorted(hp, np.arange(N), side="right")
    end = np.where(nxt < hp.size, hp[np.minimum(nxt, hp.size - 1)], N)
    dev = np.zeros(N, dtype=np.int64)
    p = si * (2 * n - si - 1) // 2 + (sj - si - 1)
    dev[p[kept]] = (C[start] + C[end] - Ns)[kept]
    return dev, Ns


@pytest.mark.parametrize("levels", [5, 60, 100000])
def test_ranking_algorithm_equals_scipy_average_ranks(levels):
    rs = np.random.RandomState(levels)
    n = 61
    N = n * (n - 1) // 2
    v = (rs.randint(0, levels, size=N) / np.float32(levels)).astype(np.float32)
    v[rs.randint(0, N, 20)] = -0.0
    iu = np.triu_indices(n, 1)
    for trial in range(5):
        mask = np.zeros(n, bool)
        mask[rs.choice(n, int(0.9 * n), replace=False)] = True
        if trial == 0:
            mask[:] = True
        dev, Ns = _device_algorithm_numpy(v, n, mask)
        kept = mask[iu[0]] & mask[iu[1]]
        assert Ns == kept.sum()
        np.testing.assert_array_equal(dev[kept], orc.rank2x(v[kept]) - (Ns + 1))
        assert not dev[~kept].any()


@pytest.mark.parametrize("n,B", [(1, 3), (2, 4), (3, 5), (10, 7), (33, 20), (64, 9), (100, 50), (1000, 40), (1854, 6)])
def test_mask_only_draw_matches_the_index_draw(n, B):
    """vr_host_draw_subsample_masks == bit masks of rng.choice(n, int(0.9 n), replace=False) draws, and
    leaves the generator exactly where numpy's loop leaves it (evals.py:356-364)."""
    from visreps_b200 import rsa as vrsa
    m = int(n * 0.9)
    ref_rng = np.random.RandomState(42)
    want = np.zeros((B, (n + 31) // 32 * 32), dtype=bool)
    for b in range(B):
        want[b, ref_rng.choice(n, size=m, replace=False)] = True
    want_words = np.packbits(want, axis=1, bitorder="little").view(np.uint32)
    rng = np.random.RandomState(42)
    got = vrsa.draw_subsample_masks(n, B, rng=rng)
    assert len(got) == B
    np.testing.assert_array_equal(got.words.numpy().view(np.uint32), want_words)
    assert rng.randint(0, 2**31 - 1) == ref_rng.randint(0, 2**31 - 1)       # generators still in step
    # a rank's share: same rows, same final generator state
    rng2 = np.random.RandomState(42)
    lo, hi = B // 3, B - B // 4
    part = vrsa.draw_subsample_masks(n, B, rng=rng2, replicate_range=(lo, hi))
    np.testing.assert_array_equal(part.words.numpy().view(np.uint32), want_words[lo:hi])
    rng3 = np.random.RandomState(42)
    vrsa.draw_subsample_masks(n, B, rng=rng3, replicate_range=(lo, hi))
    rng4 = np.random.RandomState(42)
    vrsa.draw_index_sets(n, B, rng=rng4)
    assert np.array_equal(rng3.get_state()[1], rng4.get_state()[1]) and rng3.get_state()[2] == rng4.get_state()[2]


def test_seed_keyed_draws_are_cached_and_caller_generators_are_not():
    """Draws from a bare seed are a pure function of (n, B, seed, range): the product keeps the last few (the secondary callers
    bootstrap many pairs with the same seed, run_reconstruction.py:137-158).  A generator passed by the caller is always
    replayed (its state must advance as the reference loop advances it); MASK_CACHE = False replays every call."""
    from visreps_b200 import rsa as vrsa
    vrsa._mask_cache.clear()
    a = vrsa.draw_subsample_masks(120, 9, seed=5)
    b = vrsa.draw_subsample_masks(120, 9, seed=5)
    assert a is b and a.cached
    c = vrsa.draw_subsample_masks(120, 9, seed=6)
    assert c is not a and not np.array_equal(c.words.numpy(), a.words.numpy())
    part = vrsa.draw_subsample_masks(120, 9, seed=5, replicate_range=(2, 7))
    np.testing.assert_array_equal(part.words.numpy(), a.words.numpy()[2:7])
    rng = np.random.RandomState(5)
    d = vrsa.draw_subsample_masks(120, 9, rng=rng)
    assert d is not a and not d.cached
    np.testing.assert_array_equal(d.words.numpy(), a.words.numpy())
    ref = np.random.RandomState(5)
    for _ in range(9):
        ref.choice(120, size=108, replace=False)
    assert rng.randint(0, 2**31 - 1) == ref.randint(0, 2**31 - 1)
    old = vrsa.MASK_CACHE
    try:
        vrsa.MASK_CACHE = False
        e = vrsa.draw_subsample_masks(120, 9, seed=5)
        assert e is not a and not e.cached
        np.testing.assert_array_equal(e.words.numpy(), a.words.numpy())
    finally:
        vrsa.MASK_CACHE = old
    for s_ in range(20):                                           # bounded
        vrsa.draw_subsample_masks(40, 3, seed=100 + s_)
    assert len(vrsa._mask_cache) <= vrsa._MASK_CACHE_MAX


def test_mask_only_draw_from_a_generator_in_mid_stream():
    from visreps_b200 import rsa as vrsa
    a, b = np.random.RandomState(7), np.random.RandomState(7)
    a.standard_normal(1001); b.standard_normal(1001)          # odd count: leaves a cached gaussian and a mid-block position
    a.choice(500, 77, replace=False); b.choice(500, 77, replace=False)
    got = vrsa.draw_subsample_masks(300, 25, rng=a)
    want = np.zeros((25, 320), dtype=bool)
    for r in range(25):
        want[r, b.choice(300, size=270, replace=False)] = True
    np.testing.assert_array_equal(got.words.numpy().view(np.uint32), np.packbits(want, axis=1, bitorder="little").view(np.uint32))
    assert a.standard_normal() == b.standard_normal()


@pytest.mark.parametrize("b", [1, 2, 3, 10, 40, 41, 100, 999, 1000, 1001])
def test_percentile_ci_is_numpy_percentile(b):
    """evals.py:371-372 / rsa.py:262-263 use np.percentile(scores, [2.5, 97.5]); ours must be the same bits."""
    from visreps_b200.rsa import percentile_ci
    rs = np.random.RandomState(b)
    x = rs.standard_normal((5, 4, b))
    x[0, 0, :] = 1.0
    if b > 3:
        x[1, 1, 2] = np.nan
    want = np.percentile(x, [2.5, 97.5], axis=2)
    lo, hi = percentile_ci(x)
    assert np.array_equal(want[0], lo, equal_nan=True) and np.array_equal(want[1], hi, equal_nan=True)
    lo1, hi1 = percentile_ci(x[2, 3])
    assert float(lo1) == np.percentile(x[2, 3], 2.5) and float(hi1) == np.percentile(x[2, 3], 97.5)


def test_install_rebinds_a_visreps_like_module_tree(monkeypatch):
    """install() patches the names the reference binds (evals.py:25, alignment.py:8) on whichever modules exist,
    leaves experiment scripts alone unless they are already imported, and uninstall() restores everything."""
    import sys, types
    import visreps_b200 as vb
    from visreps_b200 import evals as vevals, rsa as vrsa
    sentinel = object()
    mods = {}
    for name, syms in {"visreps": (), "visreps.analysis": (), "visreps.analysis.rsa": ("compute_rdm", "compute_rdm_correlation", "compute_rsa", "_rank"),
                       "visreps.evals": ("compute_rdm", "compute_rdm_correlation", "unrelated"),
                       "visreps.analysis.alignment": ("compute_rsa",),
                       "experiments": (), "experiments.binary_pc_rsa": (), "experiments.binary_pc_rsa.main": ("compute_rdm", "_compare_rdms")}.items():
        m = types.ModuleType(name)
        for s in syms:
            setattr(m, s, sentinel)
        mods[name] = m
        monkeypatch.setitem(sys.modules, name, m)
    done = vb.install()
    assert ("visreps.evals", "compute_rdm") in done and ("visreps.analysis.alignment", "compute_rsa") in done
    assert mods["visreps.evals"].compute_rdm is vrsa.compute_rdm and mods["visreps.evals"].unrelated is sentinel
    assert mods["visreps.analysis.rsa"]._rank is vrsa._rank
    assert mods["experiments.binary_pc_rsa.main"]._compare_rdms is vevals._compare_rdms
    assert not any(m == "experiments.reconstruction_analysis.run_reconstruction" for m, _ in done)   # not imported: not touched
    vb.uninstall()
    assert mods["visreps.evals"].compute_rdm is sentinel and mods["experiments.binary_pc_rsa.main"]._compare_rdms is sentinel


def test_evals_blocks_assemble_reference_rows_from_a_grouped_job(monkeypatch):
    """visreps_b200.evals host logic (no GPU): layers are re-extracted in sorted order (evals.py:309), every pair reads
    the column of its own best layer, rows keep the caller's (region, subject) order and the reference's keys, and the
    selection takes the first strict maximum (evals.py:270-273)."""
    import torch
    from visreps_b200 import evals as vevals, rsa as vrsa
    calls = []

    def fake_grouped(feats, neur, method, bootstrap=True, n_bootstrap=1000, seed=42, **kw):
        calls.append((len(feats), len(neur), method, bootstrap, n_bootstrap, seed))
        nA, nB = len(feats), len(neur)
        point = np.array([[10.0 * i + j for j in range(nB)] for i in range(nA)])
        out = {"point": point}
        if bootstrap:
            out["scores"] = point[:, :, None] + np.arange(n_bootstrap)[None, None, :] / 100.0
            out["ci_low"], out["ci_high"] = vrsa.percentile_ci(out["scores"])
        return out

    monkeypatch.setattr(vrsa, "rsa_grouped", fake_grouped)
    feats = {"fc1": torch.zeros(6, 3), "conv2": torch.zeros(6, 2, 2), "conv1": torch.zeros(6, 5)}
    neural = {("V1", 0): torch.zeros(6, 4), ("V1", 1): torch.zeros(6, 4), ("PPA", 0): torch.zeros(6, 2)}
    best = {("V1", 0): "fc1", ("V1", 1): "conv1", ("PPA", 0): "fc1"}
    rows = vevals.score_pairs(feats, neural, best, "Spearman", bootstrap=True, n_bootstrap=5,
                              selection_scores={k: [{"layer": "x", "score": 0.0}] for k in neural})
    assert calls[-1] == (2, 3, "spearman", True, 5, 42)              # only the distinct best layers, sorted: conv1, fc1
    assert [r["layer"] for r in rows] == ["fc1", "conv1", "fc1"]
    assert [r["score"] for r in rows] == [10.0, 1.0, 12.0]           # (fc1 = row 1, pair 0), (conv1 = row 0, pair 1), (fc1, pair 2)
    assert rows[0]["bootstrap_scores"] == [10.0, 10.01, 10.02, 10.03, 10.04] and rows[0]["compare_method"] == "spearman"
    assert rows[1]["ci_low"] == float(np.percentile(rows[1]["bootstrap_scores"], 2.5)) and rows[1]["analysis"] == "rsa"
    assert rows[2]["layer_selection_scores"] == [{"layer": "x", "score": 0.0}]
    plain = vevals.score_pairs(feats, neural, best, "pearson")
    assert all(r["ci_low"] is None and "bootstrap_scores" not in r and r["layer_selection_scores"] == [] for r in plain)

    def fake_select(feats, neur, method, bootstrap=True, **kw):
        return {"point": np.array([[0.2], [float("nan")], [0.7], [0.7]])}

    monkeypatch.setattr(vrsa, "rsa_grouped", fake_select)
    acts = {k: torch.zeros(12, 3) for k in ("a", "b", "c", "d")}
    best_layer, scores = vevals.select_layer(acts, torch.zeros(12, 2), "spearman", n_select=1000)
    assert best_layer == "c" and [s["layer"] for s in scores] == ["a", "b", "c", "d"] and np.isnan(scores[1]["score"])
    assert list(vevals.selection_indices(2000, 1000)[:3]) == list(np.random.RandomState(42).choice(2000, size=1000, replace=False)[:3])
