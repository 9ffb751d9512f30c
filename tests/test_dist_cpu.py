"""N > 1 host logic on CPU: gloo backend, world_size 2 (and a ragged 3).  The compute function is a
deterministic stand-in keyed on the replicate's index set, so the test checks exactly what the
distributed layer owns: the shard arithmetic, the shared RNG stream, and the all-gather order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from visreps_b200 import dist as vdist
from visreps_b200 import rsa as vrsa


def test_shard_bounds_partition():
    for total, world in [(1000, 8), (1001, 8), (7, 3), (3, 8), (0, 2)]:
        spans = [vdist.shard_bounds(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[k][1] == spans[k + 1][0] for k in range(world - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def _fake_compute(model_features, neural_responses, method, bootstrap, n_bootstrap, seed, replicate_range):
    n = model_features[0].shape[0]
    idx = vrsa.draw_index_sets(n, n_bootstrap, seed)[replicate_range[0]:replicate_range[1]]
    nA, nB = len(model_features), len(neural_responses)
    sc = np.empty((nA, nB, len(idx)))
    for a in range(nA):
        for b in range(nB):
            sc[a, b] = (idx.astype(np.float64) * (a + 1)).sum(axis=1) / (b + 2)   # depends on the exact draw
    return {"point": np.ones((nA, nB)), "scores": sc}


def _worker(rank, world, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        feats = [torch.zeros(40, 3), torch.zeros(40, 5)]
        neur = [torch.zeros(40, 2)]
        res = vdist.rsa_grouped_sharded(feats, neur, "spearman", n_bootstrap=total, seed=42, compute=_fake_compute)
        lo, hi = vdist.shard_bounds(total, rank, world)
        local = torch.arange(lo, hi, dtype=torch.float64).repeat(2, 1)
        order = vdist.gather_replicate_scores(local, total)
        # the blob exchange of the sharded RDM build: ragged row blocks of bytes, rank order preserved
        counts = [vdist.shard_bounds(5, r, world)[1] - vdist.shard_bounds(5, r, world)[0] for r in range(world)]
        a, b = vdist.shard_bounds(5, rank, world)
        mine = torch.stack([torch.full((16,), 10 * k + 1, dtype=torch.uint8) for k in range(a, b)]) if b > a else torch.empty((0, 16), dtype=torch.uint8)
        rows = vdist._allgather_rows(mine, counts)
        assert rows.shape == (5, 16) and [int(rows[k, 0]) for k in range(5)] == [1, 11, 21, 31, 41]
        if rank == 0:
            np.save(out, np.concatenate([res["scores"].ravel(), order.numpy().ravel()]))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,total", [(2, 10), (3, 7)])
def test_sharded_bootstrap_over_gloo_matches_single_process(world, total, tmp_path):
    out = str(tmp_path / "r0.npy")
    mp.spawn(_worker, args=(world, _free_port(), total, out), nprocs=world, join=True)
    got = np.load(out)
    feats = [torch.zeros(40, 3), torch.zeros(40, 5)]
    want = _fake_compute(feats, [torch.zeros(40, 2)], "spearman", True, total, 42, (0, total))["scores"]
    np.testing.assert_array_equal(got[:want.size], want.ravel())
    np.testing.assert_array_equal(got[want.size:], np.tile(np.arange(total, dtype=np.float64), 2))


def _np_stages():
    """CPU stand-ins with the arithmetic of the device stages (fp64 moments; fp32-shifted Gram; rsa.py:80-92 epilogue)."""
    def moments(x):
        xd = x.double()
        return torch.stack([xd.sum(1), (xd * xd).sum(1)])

    def gram(x, sums, d_total):
        shift = (sums[0] / d_total).float()
        z = x.double() - shift.double()[:, None]
        return torch.triu(z @ z.T).float(), shift

    def finish(g, sums, shift, d_total, correction):
        mean = sums[0] / d_total
        delta = mean - shift.double()
        var = (sums[1] / d_total - mean * mean).clamp_(min=0)
        sd = (var + correction).sqrt().float()
        sd[sd < correction * 10] = 1.0
        full = torch.triu(g.double()) + torch.triu(g.double(), 1).T
        cov = (full / d_total - delta[:, None] * delta[None, :]).float()
        corr = (cov / (sd[:, None] * sd[None, :] + correction)).clamp_(-1, 1)
        rdm = 1.0 - corr
        rdm.fill_diagonal_(0.0)
        return rdm, None

    return moments, gram, finish


def _ksplit_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(3)
        x = torch.from_numpy(np.maximum(rs.standard_normal((50, 301)) + 2.0 * rs.rand(50, 1), 0).astype(np.float32))
        lo, hi = vdist.column_shard(301, rank, world)
        rdm = vdist.compute_rdm_ksplit(x[:, lo:hi], stages=_np_stages())          # d_total found by an all-reduce
        rdm2 = vdist.compute_rdm_ksplit(x[:, lo:hi], d_total=301, stages=_np_stages())
        assert torch.equal(rdm, rdm2)
        many = vdist.compute_rdm_ksplit_many([x[:, lo:hi], 2.0 * x[:, lo:hi] + 1.0], [301, 301], stages=_np_stages())
        assert float((many[0] - rdm).abs().max()) <= 1e-6 and float((many[1] - rdm).abs().max()) <= 1e-5   # affine rows: same correlations
        if rank == 0:
            np.save(out, rdm.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ksplit_rdm_orchestration_over_gloo(world, tmp_path):
    """Moments all-reduce -> shared shift -> Gram all-reduce BEFORE the epilogue -> same RDM as the unsplit formula."""
    from oracle import rsa_oracle as orc
    spans = [vdist.column_shard(301, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == 301 and all(spans[k][1] == spans[k + 1][0] for k in range(world - 1))
    assert all(lo % 32 == 0 for lo, _ in spans)
    out = str(tmp_path / "rdm.npy")
    mp.spawn(_ksplit_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    rs = np.random.RandomState(3)
    x = torch.from_numpy(np.maximum(rs.standard_normal((50, 301)) + 2.0 * rs.rand(50, 1), 0).astype(np.float32))
    got = np.load(out)
    assert float(np.abs(got - orc.rdm_truth_f64(x)).max()) <= 1e-6
    assert float(np.abs(got - orc.rdm(x).numpy()).max()) <= 1e-5
