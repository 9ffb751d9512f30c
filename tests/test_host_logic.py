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
