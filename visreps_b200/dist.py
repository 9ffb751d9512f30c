"""Multi-GPU sharding of the RSA hot path: one process per GPU, torch.distributed for the plumbing.

Bootstrap replicates (and (model, layer, ROI) pairs) are independent, so ranks share
no data on the compute path.  The only exchange is one all-gather of per-replicate score vectors, in
replicate order -- the order is contractual because the reference's plotters average replicate i
element-wise across runs (plotters/plotter_utils.py:65-112).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `total` units owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_replicate_scores(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather the last axis of `local` ((..., b_local), this rank's contiguous replicate block) into
    (..., total) on every rank, replicate order preserved.  One collective; ragged shards are padded."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(total, rank, world)
    if local.shape[-1] != hi - lo:
        raise ValueError(f"rank {rank} holds {local.shape[-1]} replicates, expected {hi - lo}")
    width = -(-total // world)
    lead = local.shape[:-1]
    buf = torch.zeros(lead + (width,), dtype=local.dtype, device=local.device)
    buf[..., :hi - lo] = local
    flat = buf.contiguous().view(-1)
    out = torch.empty(world * flat.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    out = out.view((world,) + tuple(buf.shape))
    parts = []
    for r in range(world):
        a, b = shard_bounds(total, r, world)
        parts.append(out[r][..., :b - a])
    return torch.cat(parts, dim=-1)


def rsa_grouped_sharded(model_features: Sequence[torch.Tensor], neural_responses: Sequence[torch.Tensor],
                        method: str = "spearman", n_bootstrap: int = 1000, seed: int = 42, group=None,
                        compute=None) -> Dict[str, np.ndarray]:
    """`rsa_grouped` with the replicates of every pair sharded over the ranks of `group`.

    Every rank draws the same index sets from RandomState(seed) (host, cheap) and scores its own contiguous
    block; one all-gather returns all B scores per pair to every rank.  `compute` is injectable for tests.
    """
    from . import rsa
    compute = compute or rsa.rsa_grouped
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(n_bootstrap, rank, world)
    res = compute(model_features, neural_responses, method, bootstrap=True, n_bootstrap=n_bootstrap, seed=seed,
                  replicate_range=(lo, hi))
    if world == 1:
        return res
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    scores = gather_replicate_scores(torch.from_numpy(res["scores"]).to(dev), n_bootstrap, group).cpu().numpy()
    ci_low, ci_high = rsa.percentile_ci(scores)
    return {"point": res["point"], "scores": scores, "ci_low": ci_low, "ci_high": ci_high}
