"""The row-sweep Spearman engine (csrc/sweep.cu: one warp per canonical pair x 32 replicates) against the oracle and against
the legacy engine (csrc/rank.cu), which the bit-exact rank tests pin: same integers in, same scores out."""
import numpy as np
import pytest
import torch

import visreps_b200 as vb
from visreps_b200 import _lib, rsa as vrsa
from oracle import cases, rsa_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


def _both_engines(A, B, masks):
    out = {}
    for engine in ("sweep", "legacy"):
        old = vrsa.SPEARMAN_ENGINE
        vrsa.SPEARMAN_ENGINE = engine
        try:
            out[engine] = vrsa.compare_grouped(A, B, "spearman", masks).cpu().numpy()
        finally:
            vrsa.SPEARMAN_ENGINE = old
    return out["sweep"], out["legacy"]


def _rdm_from_vector(v, n, dev):
    return cases.vector_to_rdm(np.asarray(v, dtype=np.float32), n).to(dev)


@pytest.mark.parametrize("n,levels,R", [(40, 5, 7), (130, 300, 33), (300, 40, 64), (700, 100000, 70)])
def test_sweep_equals_legacy_and_oracle_with_ties(n, levels, R, dev):
    """Tie structures from 'almost everything tied' to 'almost nothing tied'; replicate counts around the 32-lane batches."""
    N = n * (n - 1) // 2
    a, b = cases.tied_vectors(N, levels, 900 + n)
    rs = np.random.RandomState(n)
    c = rs.standard_normal(N).astype(np.float32)
    A = vrsa.PresortedRDM.batch([_rdm_from_vector(a, n, dev), _rdm_from_vector(c, n, dev), _rdm_from_vector(np.round(c, 1), n, dev)])
    B = vrsa.PresortedRDM.batch([_rdm_from_vector(b, n, dev), _rdm_from_vector(-c + 0.01 * a, n, dev)])
    idx = vrsa.draw_index_sets(n, R - 1, 42)
    masks = vrsa.build_masks(idx, n, dev, with_full=True)
    sweep, legacy = _both_engines(A, B, masks)
    assert sweep.shape == (3, 2, R)
    np.testing.assert_array_equal(sweep, legacy)                      # exact integer sums on both sides: bit-identical
    for r in (0, 1, R - 1):                                           # point estimate and two replicates against SciPy
        sub = np.arange(n) if r == 0 else np.sort(idx[r - 1])
        for i, va in enumerate((a, c, np.round(c, 1))):
            ra = cases.vector_to_rdm(np.asarray(va, dtype=np.float32), n)[sub][:, sub]
            rb = cases.vector_to_rdm(b, n)[sub][:, sub]
            want = orc.rdm_compare(ra, rb, correlation="Spearman")
            assert abs(sweep[i, 0, r] - want) <= 1e-9


def test_sweep_run_shapes_long_runs_cross_group_and_constant(dev):
    """Runs that end in the next 32-group, runs longer than a group, one run spanning the whole vector, runs at the very end."""
    n = 66
    N = n * (n - 1) // 2                                              # 2145 = 67 groups + 1 position
    base = np.arange(N, dtype=np.float32)
    v1 = base.copy(); v1[20:45] = 20.0; v1[60:70] = 60.0; v1[100:400] = 100.0; v1[N - 3:] = float(N)      # cross, long, tail runs
    v2 = np.floor(base / 31.0).astype(np.float32)                      # every run crosses a group boundary somewhere
    v3 = np.full(N, 3.0, dtype=np.float32)                             # constant: correlation undefined -> NaN
    v4 = base[::-1].copy()
    rs = np.random.RandomState(1)
    perm = rs.permutation(N)
    vecs = [v1[perm], v2[perm], v3, v4[perm]]
    A = vrsa.PresortedRDM.batch([_rdm_from_vector(v, n, dev) for v in vecs])
    masks = vrsa.build_masks(vrsa.draw_index_sets(n, 40, 7), n, dev, with_full=True)
    sweep, legacy = _both_engines(A, A, masks)
    assert np.array_equal(sweep, legacy, equal_nan=True)
    assert np.isnan(sweep[2]).all() and np.isnan(sweep[:, 2]).all()
    for k in (0, 1, 3):
        np.testing.assert_allclose(sweep[k, k], 1.0, atol=1e-12, rtol=0)
    want = orc.rdm_compare(cases.vector_to_rdm(vecs[0], n), cases.vector_to_rdm(vecs[1], n), correlation="Spearman")
    assert abs(sweep[0, 1, 0] - want) <= 1e-9 and sweep[0, 1, 0] == sweep[1, 0, 0]


@pytest.mark.parametrize("n,R", [(66, 40), (66, 256), (300, 70), (2100, 40)])
def test_one_pair_jobs_take_the_batch_inner_tables(n, R, dev):
    """One vector against one vector with 33-256 masks: rows are laid out batch-innermost and `sweep_pair_kernel` walks all
    batches of a pair (sweep.cu: sweep_batch_inner).  Every run shape (single, inside a group, ending in the next group, longer
    than a group, longer than the run word holds, at the very end), compact (n <= 2048) and general run words; same bits as
    the legacy engine, which never sees that layout."""
    N = n * (n - 1) // 2
    rs = np.random.RandomState(n + R)
    base = np.arange(N, dtype=np.float32)
    v1 = base.copy(); v1[20:45] = 20.0; v1[60:70] = 60.0; v1[100:400] = 100.0; v1[N - 3:] = float(N)
    v2 = np.floor(base / 31.0).astype(np.float32)
    v3 = rs.standard_normal(N).astype(np.float32)
    v4 = np.round(rs.standard_normal(N), 2 if n < 1000 else 3).astype(np.float32)
    perm = rs.permutation(N)
    vecs = [v1[perm], v2[rs.permutation(N)], v3, v4]
    pres = [vrsa.PresortedRDM(_rdm_from_vector(v, n, dev)) for v in vecs]
    masks = vrsa.build_masks(vrsa.draw_index_sets(n, R - 1, 11), n, dev, with_full=True)
    for i, j in ((0, 1), (1, 2), (2, 3), (3, 0)):
        sweep, legacy = _both_engines([pres[i]], [pres[j]], masks)
        assert sweep.shape == (1, 1, R)
        if N ** 3 < 9e18:
            np.testing.assert_array_equal(sweep, legacy)
        else:
            np.testing.assert_allclose(sweep, legacy, atol=1e-12, rtol=0)
    want = orc.rdm_compare(cases.vector_to_rdm(vecs[0], n), cases.vector_to_rdm(vecs[1], n), correlation="Spearman")
    got = vrsa.compare_grouped([pres[0]], [pres[1]], "spearman", masks).cpu().numpy()
    assert abs(got[0, 0, 0] - want) <= 1e-9


def test_sweep_tiles_and_group_sizes_agree(dev):
    """Any tiling of the (A, B) grid (options sweep_tile_a / sweep_tile_b, padded instantiations for 3, 5, 6, 7 vectors) gives
    the same bits; more than eight vectors per side are tiled by the library."""
    x, y = cases.pair_inputs("p220")
    n = x.shape[0]
    rs = np.random.RandomState(3)
    feats = [x[:, rs.permutation(x.shape[1])[:40 + 3 * k]] for k in range(11)]
    neur = [y[:, rs.permutation(y.shape[1])[:30 + 5 * k]] for k in range(5)]
    A = vrsa.PresortedRDM.batch([vb.compute_rdm(f.to(dev)) for f in feats])
    B = vrsa.PresortedRDM.batch([vb.compute_rdm(f.to(dev)) for f in neur])
    masks = vrsa.build_masks(vrsa.draw_index_sets(n, 45, 42), n, dev, with_full=True)
    ref, legacy = _both_engines(A, B, masks)
    np.testing.assert_array_equal(ref, legacy)
    try:
        for ta, tb in ((1, 1), (3, 5), (8, 2), (5, 8), (7, 3)):
            _lib.set_option("sweep_tile_a", ta)
            _lib.set_option("sweep_tile_b", tb)
            got = vrsa.compare_grouped(A, B, "spearman", masks).cpu().numpy()
            np.testing.assert_array_equal(got, ref)
    finally:
        _lib.set_option("sweep_tile_a", 8)
        _lib.set_option("sweep_tile_b", 8)
    with pytest.raises(ValueError):
        _lib.set_option("sweep_tile_a", 9)


def test_sweep_large_n_general_run_words_and_fp64_sums(dev):
    """n = 2100: N = 2,203,950 pairs is past the compact run words (n <= 2048) and past exact int64 totals; the sweep then
    sums in fp64 in a fixed order.  Checked against the legacy engine (its own fp64 path) and SciPy on the point estimate."""
    n = 2100
    rs = np.random.RandomState(5)
    z = rs.standard_normal((n, 12)).astype(np.float32)
    x = torch.from_numpy(np.maximum(z @ rs.standard_normal((12, 60)).astype(np.float32) + rs.standard_normal((n, 60)).astype(np.float32), 0))
    y = torch.from_numpy(np.round(z @ rs.standard_normal((12, 9)).astype(np.float32), 1))           # coarse values: many ties
    A = vrsa.PresortedRDM.batch([vb.compute_rdm(x.to(dev))])
    B = vrsa.PresortedRDM.batch([vb.compute_rdm(y.to(dev))])
    masks = vrsa.build_masks(vrsa.draw_index_sets(n, 3, 42), n, dev, with_full=True)
    sweep, legacy = _both_engines(A, B, masks)
    np.testing.assert_allclose(sweep, legacy, atol=1e-12, rtol=0)
    again = vrsa.compare_grouped(A, B, "spearman", masks).cpu().numpy()
    np.testing.assert_array_equal(sweep, again)                        # deterministic
    want = orc.rdm_compare(A[0].rdm.cpu(), B[0].rdm.cpu(), correlation="Spearman")
    assert abs(sweep[0, 0, 0] - want) <= 1e-9


def test_sweep_job_is_incremental(dev):
    """Vectors added to a job one by one, in any order, and swept in groups give the scores of the one-shot call."""
    x, y = cases.pair_inputs("p64")
    n = x.shape[0]
    rdms = [vb.compute_rdm(x[:, k::3].to(dev)) for k in range(3)] + [vb.compute_rdm(y[:, k::2].to(dev)) for k in range(2)]
    items = vrsa.PresortedRDM.batch(rdms)
    masks = vrsa.build_masks(vrsa.draw_index_sets(n, 20, 42), n, dev, with_full=True)
    R = int(masks.shape[0])
    want = vrsa.compare_grouped(items[:3], items[3:], "spearman", masks)
    job = vrsa.SpearmanJob(masks, n, 5)
    job.add([items[4]], 4)
    job.add([items[3]], 3)
    out = torch.full((3, 2, R), float("nan"), dtype=torch.float64, device=dev)
    for k in (2, 0, 1):
        job.add([items[k]], k)
        job.sweep([k], [3, 4], out.view(-1), k * 2 * R, 2 * R, R)
    assert torch.equal(out, want)
    with pytest.raises(ValueError):
        job.add([vrsa.PresortedRDM(vb.compute_rdm(torch.randn(n + 1, 8, device=dev)))], 0)


def test_nan_entries_make_only_the_replicates_that_keep_them_nan(dev):
    """rsa.py:123-125: a replicate is NaN when ITS sub-RDM holds a NaN; replicates that drop the offending stimulus are
    finite and equal the oracle.  Same rule in both Spearman engines; +NaN and -NaN entries (sorted to opposite ends)."""
    x, y = cases.pair_inputs("p64")
    n = x.shape[0]
    a, b = orc.rdm(x), orc.rdm(y)
    bad = a.clone()
    bad[5, :] = float("nan"); bad[:, 5] = float("nan"); bad[5, 5] = 0.0          # stimulus 5 is NaN everywhere
    neg = torch.tensor(float("nan")).view(torch.int32).bitwise_or(torch.tensor(-2**31, dtype=torch.int32)).view(torch.float32)
    bad[9, 20] = neg; bad[20, 9] = neg                                          # one -NaN pair (stimuli 9 and 20)
    idx = orc.draw_subsamples(n, 60, 42)
    masks = vrsa.build_masks(idx, n, dev, with_full=True)
    A = vrsa.PresortedRDM.batch([bad.to(dev), a.to(dev)])
    B = vrsa.PresortedRDM.batch([b.to(dev)])
    sweep, legacy = _both_engines(A, B, masks)
    assert np.array_equal(sweep, legacy, equal_nan=True)
    assert np.isnan(sweep[0, 0, 0]) and np.isfinite(sweep[1, 0]).all()
    n_clean = 0
    for r in range(60):
        kept = set(idx[r].tolist())
        dirty = 5 in kept or (9 in kept and 20 in kept)
        assert np.isnan(sweep[0, 0, r + 1]) == dirty, r
        if not dirty:
            n_clean += 1
            sub = torch.from_numpy(idx[r].astype(np.int64))
            want = orc.rdm_compare(bad[sub][:, sub], b[sub][:, sub], correlation="Spearman")
            assert abs(sweep[0, 0, r + 1] - want) <= 1e-9
    assert n_clean >= 1


def test_rsa_grouped_host_pipeline_through_the_sweep_engine(dev, monkeypatch):
    """Host inputs whose count tables exceed shared memory (n > 1277), or any host input once the legacy threshold is lowered,
    are added to the job one matrix at a time as their copies land and swept against the other side: same bits as the batched
    device-resident job and as the legacy pipeline."""
    rs = np.random.RandomState(23)
    n, B = 140, 37
    z = rs.standard_normal((n, 10)).astype(np.float32)
    feats = [torch.from_numpy(np.maximum(z @ rs.standard_normal((10, d)).astype(np.float32) + rs.standard_normal((n, d)).astype(np.float32), 0))
             for d in (300, 40, 1200, 77, 513)]
    neur = [torch.from_numpy(np.round(z @ rs.standard_normal((10, v)) + rs.standard_normal((n, v)), 2).astype(np.float32)) for v in (25, 60, 33)]
    on_device = vb.rsa_grouped([f.to(dev) for f in feats], [y.to(dev) for y in neur], "spearman", n_bootstrap=B, seed=42)
    legacy_host = vb.rsa_grouped([f.pin_memory() for f in feats], neur, "spearman", n_bootstrap=B, seed=42)
    monkeypatch.setattr(vrsa, "_LEGACY_SMEM_TABLE_MAX_N", 0)                 # force the sweep pipeline for host inputs
    sweep_host = vb.rsa_grouped([f.pin_memory() for f in feats], neur, "spearman", n_bootstrap=B, seed=42)
    swapped = vb.rsa_grouped(neur, [f.pin_memory() for f in feats], "spearman", n_bootstrap=B, seed=42)
    for key in ("point", "scores", "ci_low", "ci_high"):
        np.testing.assert_array_equal(sweep_host[key], on_device[key])
        np.testing.assert_array_equal(legacy_host[key], on_device[key])
    np.testing.assert_array_equal(swapped["scores"], np.transpose(on_device["scores"], (1, 0, 2)))


def test_kendall_complement_mode_equals_the_direct_count(dev):
    """csrc/kendall.cu: ties and inversions counted among the DROPPED entries (complement mode, default) give exactly the tau-a
    of the direct count among the kept ones, for masks that keep most, half, few and all stimuli, with ties, and equal SciPy."""
    rs = np.random.RandomState(9)
    n = 90
    N = n * (n - 1) // 2
    a, b = cases.tied_vectors(N, 60, 77)
    ra, rb = cases.vector_to_rdm(a, n).to(dev), cases.vector_to_rdm((b + 0.3 * rs.standard_normal(N)).astype(np.float32).round(1), n).to(dev)
    sets = [rs.choice(n, size=m, replace=False) for m in (81, 81, 81, 60, 45, 20, 3, 2, 89, 90)]
    W = (n + 31) // 32
    words = np.zeros((len(sets) + 1, W), dtype=np.uint32)
    words[0] = vrsa.full_mask(n)
    for r, idx in enumerate(sets):
        for i in idx:
            words[r + 1, i >> 5] |= np.uint32(1) << np.uint32(i & 31)
    masks = torch.from_numpy(words.view(np.int32)).to(dev)
    A, B = [vrsa._RawRDM(ra)], [vrsa._RawRDM(rb)]
    out = {}
    for mode in (1, 0, "tree"):
        _lib.set_option("kendall_tree", 2 if mode == "tree" else 0)
        _lib.set_option("kendall_complement", 0 if mode == "tree" else mode)
        lib = _lib.load()
        ws = _lib.workspace(lib.vr_kendall_batch_workspace_bytes(n, masks.shape[0]), dev)
        res = torch.empty(masks.shape[0], dtype=torch.float64, device=dev)
        _lib.check(lib.vr_kendall_tau_a(ra.data_ptr(), rb.data_ptr(), n, n, masks.data_ptr(), int(masks.shape[0]), res.data_ptr(),
                                        ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "vr_kendall_tau_a")
        out[mode] = res.cpu().numpy()
    _lib.set_option("kendall_complement", 1)
    _lib.set_option("kendall_tree", 1)
    np.testing.assert_array_equal(out[1], out[0])
    np.testing.assert_array_equal(out["tree"], out[0])                 # integer counts in every mode: same bits
    va, vb_ = orc.triu_vector(ra), orc.triu_vector(rb)
    iu = np.triu_indices(n, 1)
    for r, idx in enumerate([np.arange(n)] + sets):
        keep = np.zeros(n, dtype=bool); keep[idx] = True
        kp = keep[iu[0]] & keep[iu[1]]
        want = orc.kendall_tau_a(va[kp], vb_[kp])[0]
        got = out[1][r]
        assert (np.isnan(want) and np.isnan(got)) or abs(got - want) <= 1e-6, (r, got, want)      # the oracle goes through tau-b in floating point
    # through the public API, both modes chosen by the kept fraction
    big = vb.bootstrap_rdm_correlation(ra, rb, "kendall", n_bootstrap=7, seed=42)
    few = vb.bootstrap_rdm_correlation_grouped([ra], [rb], "kendall", index_sets=np.stack([rs.choice(n, size=30, replace=False) for _ in range(5)]))
    assert len(big[3]) == 7 and few["scores"].shape == (1, 1, 5) and np.isfinite(few["scores"]).all()


@pytest.mark.parametrize("n,R,levels", [(20, 5, 3), (66, 40, 7), (66, 300, 9), (300, 70, 25), (300, 70, 100000), (700, 200, 400)])
def test_kendall_tree_mode_counts_discordant_pairs_without_sorting(n, R, levels, dev):
    """csrc/kendall.cu tree mode: the merge tree of the RDM pair is built once, replicates ride in the lanes.  Same bits as the
    sorting path for tiny trees (no merge level), nodes inside one 2048-position segment, nodes stitched from several segments,
    heavy ties in both vectors, more than 64 replicates (two lane batches per pass) and the automatic choice (R >= 192)."""
    N = n * (n - 1) // 2
    rs = np.random.RandomState(n + R)
    a, b = cases.tied_vectors(N, levels, 31 + n)
    b = (b + 0.5 * rs.standard_normal(N).astype(np.float32)).round(2 if levels < 1000 else 5).astype(np.float32)
    ra, rb = cases.vector_to_rdm(a, n).to(dev), cases.vector_to_rdm(b, n).to(dev)
    masks = vrsa.build_masks(vrsa.draw_index_sets(n, R - 1, 5), n, dev, with_full=True)
    lib = _lib.load()
    out = {}
    try:
        for mode in (0, 2, 1):
            _lib.set_option("kendall_tree", mode)
            ws = _lib.workspace(lib.vr_kendall_batch_workspace_bytes(n, R), dev)
            res = torch.full((R,), -7.0, dtype=torch.float64, device=dev)
            l0 = _lib.launch_count()
            _lib.check(lib.vr_kendall_tau_a(ra.data_ptr(), rb.data_ptr(), n, n, masks.data_ptr(), R, res.data_ptr(),
                                            ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "vr_kendall_tau_a")
            out[mode] = res.cpu().numpy()
    finally:
        _lib.set_option("kendall_tree", 1)
    np.testing.assert_array_equal(out[2], out[0])
    np.testing.assert_array_equal(out[1], out[0])
    va, vb_ = orc.triu_vector(ra), orc.triu_vector(rb)
    want = orc.kendall_tau_a(va, vb_)[0]
    assert abs(out[2][0] - want) <= 1e-6


def test_kendall_tree_mode_nan_constant_and_tiny_subsamples(dev):
    """Tree mode against the sorting path where tau is undefined or degenerate: a constant vector (NaN, rsa.py:29-30), a NaN entry
    (NaN for every mask), masks that keep 0, 1, 2 and 3 stimuli (fewer than two kept entries -> NaN), and the full mask."""
    n = 70
    N = n * (n - 1) // 2
    rs = np.random.RandomState(4)
    a = rs.standard_normal(N).astype(np.float32).round(1)
    b = rs.standard_normal(N).astype(np.float32)
    const = np.full(N, 2.5, dtype=np.float32)
    with_nan = b.copy(); with_nan[17] = np.nan
    sets = [[], [5], [5, 9], [5, 9, 40], list(range(n)), list(range(0, n, 2))] + [list(rs.choice(n, size=63, replace=False)) for _ in range(8)]
    words = np.zeros((len(sets), (n + 31) // 32), dtype=np.uint32)
    for r, idx in enumerate(sets):
        for i in idx:
            words[r, i >> 5] |= np.uint32(1) << np.uint32(i & 31)
    masks = torch.from_numpy(words.view(np.int32)).to(dev)
    lib = _lib.load()
    R = len(sets)
    try:
        for va, vb_ in ((a, b), (const, b), (a, const), (a, with_nan)):
            ra, rb = cases.vector_to_rdm(va, n).to(dev), cases.vector_to_rdm(vb_, n).to(dev)
            out = {}
            for mode in (0, 2):
                _lib.set_option("kendall_tree", mode)
                ws = _lib.workspace(lib.vr_kendall_batch_workspace_bytes(n, R), dev)
                res = torch.full((R,), -7.0, dtype=torch.float64, device=dev)
                _lib.check(lib.vr_kendall_tau_a(ra.data_ptr(), rb.data_ptr(), n, n, masks.data_ptr(), R, res.data_ptr(),
                                                ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "vr_kendall_tau_a")
                out[mode] = res.cpu().numpy()
            np.testing.assert_array_equal(out[2], out[0])
            assert np.isnan(out[2][:3]).all()                       # 0, 1 or 2 stimuli: fewer than two kept entries
            if va is const or vb_ is const or vb_ is with_nan:
                assert np.isnan(out[2]).all()
            else:
                assert np.isfinite(out[2][3:]).all()
    finally:
        _lib.set_option("kendall_tree", 1)


@pytest.mark.parametrize("b", [1, 2, 3, 10, 40, 41, 100, 999, 1000, 1001, 5000, 16384])
def test_device_percentile_ci_is_numpy_percentile(b, dev):
    """vr_percentile_ci_f64 == np.percentile(scores, [2.5, 97.5], axis=-1) bit for bit (evals.py:371-372, rsa.py:262-263):
    negative values, ties, -0.0, infinities; a row with a NaN gives NaN bounds; strided rows (the [..., 1:] slice rsa_grouped
    passes); other percentiles."""
    rs = np.random.RandomState(b)
    x = rs.standard_normal((3, 4, b + 1)) * 0.3
    x[0, 1] = np.round(x[0, 1], 1)                                   # ties
    x[1, 2, ::3] = -0.0
    if b > 4:
        x[2, 0, 2] = np.inf; x[2, 0, 3] = -np.inf
        x[2, 3, b // 2] = np.nan
    t = torch.from_numpy(x).to(dev)
    got = vrsa.percentile_ci_device(t[:, :, 1:])
    assert got is not None and got.shape == (2, 3, 4)
    with np.errstate(invalid="ignore"):
        want = np.percentile(x[:, :, 1:], [2.5, 97.5], axis=-1)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    lo_hi = vrsa.percentile_ci_device(t[:, :, 1:], 0.0, 100.0).cpu().numpy()
    with np.errstate(invalid="ignore"):
        np.testing.assert_array_equal(lo_hi, np.percentile(x[:, :, 1:], [0.0, 100.0], axis=-1))
    mid = vrsa.percentile_ci_device(t[:, :, 1:], 50.0, 33.3).cpu().numpy()
    with np.errstate(invalid="ignore"):
        np.testing.assert_array_equal(mid, np.percentile(x[:, :, 1:], [50.0, 33.3], axis=-1))
    assert vrsa.percentile_ci_device(t[:, :, 1:].float()) is None     # not fp64: the caller keeps the host routine
    with np.errstate(invalid="ignore"):
        l, h = vrsa.percentile_ci(x[:, :, 1:])
    np.testing.assert_array_equal(np.stack([l, h]), want)              # and the host routine agrees with both
