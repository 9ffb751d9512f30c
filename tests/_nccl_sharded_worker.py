"""Worker of tests/test_gpu_dist_nccl.py: one process per GPU (torchrun), NCCL.  Runs dist.rsa_grouped_sharded with the
sharded RDM build (per-rank RDM + presort, blob all-gather, PresortedRDM.from_blob) and compares every rank's result
with the single-GPU rsa_grouped of the same inputs, bit for bit."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import visreps_b200 as vb
from visreps_b200 import dist as vdist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        rs = np.random.RandomState(7)
        n, B = 300, 37                                      # 37 replicates: ragged shards at any world size > 1
        z = rs.standard_normal((n, 12)).astype(np.float32)
        feats = [torch.from_numpy(np.maximum(z @ rs.standard_normal((12, d)).astype(np.float32)
                                             + rs.standard_normal((n, d)).astype(np.float32), 0)) for d in (40, 333, 64)]
        neur = [torch.from_numpy((z @ rs.standard_normal((12, v)) + rs.standard_normal((n, v))).astype(np.float32)) for v in (50, 21)]
        feats = [f.to(dev) for f in feats]
        neur = [q.to(dev) for q in neur]
        want = vb.rsa_grouped(feats, neur, "spearman", bootstrap=True, n_bootstrap=B, seed=42)
        got = vdist.rsa_grouped_sharded(feats, neur, "spearman", n_bootstrap=B, seed=42, shard_rdms="force")
        for key in ("point", "scores", "ci_low", "ci_high"):
            assert np.array_equal(got[key], want[key]), (rank, key, np.abs(got[key] - want[key]).max())
        # replicate-only sharding (every rank builds every RDM) agrees as well
        got2 = vdist.rsa_grouped_sharded(feats, neur, "spearman", n_bootstrap=B, seed=42, shard_rdms=False)
        assert np.array_equal(got2["scores"], want["scores"]) and np.array_equal(got2["point"], want["point"])
        got3 = vdist.rsa_grouped_sharded(feats, neur, "pearson", n_bootstrap=B, seed=42)
        want3 = vb.rsa_grouped(feats, neur, "pearson", bootstrap=True, n_bootstrap=B, seed=42)
        assert np.array_equal(got3["scores"], want3["scores"])
        # the balanced build: wide inputs K-split over the ranks inside the sharded job; scores within the 1e-5 budget of one GPU,
        # bit-identical with ksplit=False
        old_min = vdist.KSPLIT_MIN_COLS
        vdist.KSPLIT_MIN_COLS = 64
        try:
            wide_feats = [torch.cat([f] * 3, dim=1) for f in feats]                 # 120 / 999 / 192 columns
            want_w = vb.rsa_grouped(wide_feats, neur, "spearman", bootstrap=True, n_bootstrap=B, seed=42)
            got_w = vdist.rsa_grouped_sharded(wide_feats, neur, "spearman", n_bootstrap=B, seed=42, shard_rdms="force", ksplit=True)
            assert float(np.abs(got_w["scores"] - want_w["scores"]).max()) <= 1e-5 and float(np.abs(got_w["point"] - want_w["point"]).max()) <= 1e-5
            got_x = vdist.rsa_grouped_sharded(wide_feats, neur, "spearman", n_bootstrap=B, seed=42, shard_rdms="force", ksplit=False)
            assert np.array_equal(got_x["scores"], want_w["scores"])
        finally:
            vdist.KSPLIT_MIN_COLS = old_min
        # K-split RDM: every rank runs the SYRK on its column shard, Grams reduced over NCCL before the epilogue
        from oracle import rsa_oracle as orc
        for (nk, dk, seed) in ((130, 700, 1), (500, 40000, 2)):
            rk = np.random.RandomState(seed)
            xk = torch.from_numpy(np.maximum(rk.standard_normal((nk, dk)).astype(np.float32) + 3.0 * rk.rand(nk, 1).astype(np.float32), 0))
            got_k = vdist.compute_rdm_sharded(xk.to(dev))
            one = vb.compute_rdm(xk.to(dev))
            assert torch.equal(got_k, got_k.T) and float(got_k.diagonal().abs().max()) == 0.0
            assert float((got_k - one).abs().max()) <= 1e-5, float((got_k - one).abs().max())
            assert float(np.abs(got_k.cpu().numpy().astype(np.float64) - orc.rdm_truth_f64(xk)).max()) <= 1e-5
        dist.barrier()
        if rank == 0:
            print(f"nccl sharded build ok: world={world}", flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
