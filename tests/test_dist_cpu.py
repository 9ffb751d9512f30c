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
