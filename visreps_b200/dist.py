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


def _allgather_rows(mine: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """All-gather row blocks of unequal height: rank r contributes counts[r] rows of `mine` ((rows, width))."""
    world = dist.get_world_size(group)
    per = max(counts)
    buf = torch.zeros((per,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
    buf[:mine.shape[0]] = mine
    out = torch.empty((world * per,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(out.view(-1), buf.view(-1), group=group)
    out = out.view((world, per) + tuple(mine.shape[1:]))
    return torch.cat([out[r, :counts[r]] for r in range(world)], dim=0)


KSPLIT_MIN_COLS = 4096      # an input is K-split over the ranks when every rank gets at least this many columns


def rsa_grouped_sharded(model_features: Sequence[torch.Tensor], neural_responses: Sequence[torch.Tensor],
                        method: str = "spearman", n_bootstrap: int = 1000, seed: int = 42, group=None,
                        compute=None, shard_rdms: bool = True, ksplit: bool = True) -> Dict[str, np.ndarray]:
    """`rsa_grouped` with the replicates of every pair sharded over the ranks of `group`.

    Every rank draws the same index sets from RandomState(seed) (host, cheap) and scores its own contiguous
    block; one all-gather returns all B scores per pair to every rank.

    With `shard_rdms` (NCCL, Spearman) the replicate-independent work is sharded too, so that it does not cap the
    scaling: rank r builds and presorts only its share of the RDMs, and one all-gather over NVLink hands every
    rank all presorted blobs (8 bytes per RDM pair, 14 MB at n = 1854) before the replicates are ranked.
    With `ksplit` the RDMs of wide inputs (>= KSPLIT_MIN_COLS columns per rank) are built cooperatively (feature-sharded
    SYRK + all-reduce of the Gram): balanced whatever the layer widths, at the price of RDM entries that differ from the
    single-GPU build in the last bits (different summation order; scores agree to ~1e-7).  `ksplit=False` keeps every RDM on
    one rank and the result bit-identical to one GPU.
    `compute` is injectable for tests (and selects the replicate-only sharding).  `shard_rdms="force"` takes the
    sharded-build branch at any world size (a one-rank NCCL group exercises the blob exchange on one GPU).
    """
    from . import rsa
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(n_bootstrap, rank, world)
    method = method.lower()
    sharded_build = (compute is None and bool(shard_rdms) and (world > 1 or shard_rdms == "force") and method == "spearman"
                     and dist.is_initialized() and dist.get_backend(group) == "nccl")
    if not sharded_build:
        compute = compute or rsa.rsa_grouped
        res = compute(model_features, neural_responses, method, bootstrap=True, n_bootstrap=n_bootstrap, seed=seed,
                      replicate_range=(lo, hi))
        if world == 1:
            return res
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        scores = gather_replicate_scores(torch.from_numpy(res["scores"]).to(dev), n_bootstrap, group).cpu().numpy()
        ci_low, ci_high = rsa.percentile_ci(scores)
        return {"point": res["point"], "scores": scores, "ci_low": ci_low, "ci_high": ci_high}

    from . import _lib
    lib = _lib.load()
    dev = _lib.require_cuda()
    inputs = [rsa._flat(t) for t in list(model_features) + list(neural_responses)]
    nA, total = len(model_features), len(inputs)
    n = int(inputs[0].shape[0])
    if any(int(t.shape[0]) != n for t in inputs):
        raise ValueError("all inputs must have the same number of stimuli")
    # 1. RDM builds.  Wide inputs are K-split: every rank runs the tensor-core SYRK on its column shard and the Grams are
    #    all-reduced over NVLink (compute_rdm_ksplit) -- perfectly balanced whatever the layer widths are, where one RDM per
    #    rank leaves the rank that drew conv1 (39 % of all the columns of an AlexNet) as the straggler.  Narrow inputs are
    #    built whole by one rank each.  Every RDM is presorted by exactly one rank.
    flags = torch.zeros(total, dtype=torch.int32, device=dev)
    wide = [k for k in range(total) if ksplit and world > 1 and int(inputs[k].shape[1]) >= KSPLIT_MIN_COLS * world]
    narrow = [k for k in range(total) if k not in wide]
    owner = {}
    for pos, k in enumerate(narrow):
        owner[k] = pos % world
    load = [sum(1 for k in narrow if owner[k] == r) for r in range(world)]
    for k in wide:                                            # presort duty of the K-split RDMs: least loaded rank first
        r = min(range(world), key=lambda q: (load[q], q))
        owner[k] = r
        load[r] += 1
    rdms = {}
    if wide:
        cols = [column_shard(int(inputs[k].shape[1]), rank, world) for k in wide]
        fulls = compute_rdm_ksplit_many([inputs[k][:, c0:c1] for k, (c0, c1) in zip(wide, cols)],
                                        [int(inputs[k].shape[1]) for k in wide], group=group)
        for k, full in zip(wide, fulls):
            if owner[k] == rank:
                rdms[k] = full
    for k in narrow:
        if owner[k] == rank:
            rdms[k] = rsa._rdm_device(inputs[k], dev, 1e-12, flags, k)
    mine_idx = sorted(rdms)
    mine = rsa.PresortedRDM.batch([rdms[k] for k in mine_idx])
    blob_bytes = int(lib.vr_presort_bytes(n))
    local = torch.stack([it.blob for it in mine]) if mine else torch.empty((0, blob_bytes), dtype=torch.uint8, device=dev)
    # 2. the exchange step of the sharded build: every rank receives every presorted blob (rank-major order)
    counts = [sum(1 for k in range(total) if owner[k] == r) for r in range(world)]
    blobs = _allgather_rows(local, counts, group)
    order = [k for r in range(world) for k in range(total) if owner[k] == r]
    items = [None] * total
    for pos, k in enumerate(order):
        items[k] = rsa.PresortedRDM.from_blob(blobs[pos], n)
    # 3. this rank's replicates (plus the point estimate, one extra mask) for all pairs
    draws = rsa.draw_subsample_masks(n, n_bootstrap, seed, replicate_range=(lo, hi))
    masks = rsa.build_masks(draws, n, dev, with_full=True)
    out = rsa.compare_grouped(items[:nA], items[nA:], "spearman", masks)
    gathered = gather_replicate_scores(out[:, :, 1:].contiguous(), n_bootstrap, group)
    ci = rsa.percentile_ci_device(gathered) if gathered.is_cuda else None
    scores = gathered.cpu().numpy()
    if ci is not None:
        ci_low, ci_high = ci[0].cpu().numpy(), ci[1].cpu().numpy()
    else:
        ci_low, ci_high = rsa.percentile_ci(scores)
    return {"point": out[:, :, 0].cpu().numpy(), "scores": scores, "ci_low": ci_low, "ci_high": ci_high}


# --------------------------------------------------------------------------------------------------
# K-split multi-GPU RDM (BASELINE config 5: 10,000 stimuli x 193,600 features)
# --------------------------------------------------------------------------------------------------
def column_shard(d: int, rank: int, world: int, align: int = 32) -> Tuple[int, int]:
    """Feature columns [lo, hi) of rank `rank`: contiguous, multiples of `align` columns (one TMA box row) except the tail."""
    blocks = -(-d // align)
    lo, hi = shard_bounds(blocks, rank, world)
    return min(d, lo * align), min(d, hi * align)


def _rdm_ksplit_stages():
    """The four device stages of the K-split RDM (C ABI: vr_rdm_moments_f32, vr_rdm_shift_from_moments, vr_rdm_gram_f32,
    vr_rdm_finish_f32) as Python callables; injectable so the orchestration can be exercised on CPU with gloo."""
    from . import _lib

    def moments(x):
        lib = _lib.load()
        n, d = int(x.shape[0]), int(x.shape[1])
        out = torch.empty((2, n), dtype=torch.float64, device=x.device)
        _lib.check(lib.vr_rdm_moments_f32(x.data_ptr(), n, d, x.stride(0), out[0].data_ptr(), out[1].data_ptr(), _lib.stream_ptr()),
                   "vr_rdm_moments_f32")
        return out

    def gram(x, sums, d_total):
        lib = _lib.load()
        n, d = int(x.shape[0]), int(x.shape[1])
        shift = torch.empty(n, dtype=torch.float32, device=x.device)
        _lib.check(lib.vr_rdm_shift_from_moments(sums[0].data_ptr(), n, int(d_total), shift.data_ptr(), _lib.stream_ptr()),
                   "vr_rdm_shift_from_moments")
        g = torch.empty((n, n), dtype=torch.float32, device=x.device)
        ws = _lib.workspace(lib.vr_rdm_workspace_bytes(n, d), x.device)
        _lib.check(lib.vr_rdm_gram_f32(x.data_ptr(), n, d, x.stride(0), shift.data_ptr(), g.data_ptr(), n, ws.data_ptr(), ws.numel(),
                                       _lib.stream_ptr()), "vr_rdm_gram_f32")
        return g, shift

    def finish(g, sums, shift, d_total, correction):
        lib = _lib.load()
        n = int(g.shape[0])
        rdm = torch.empty((n, n), dtype=torch.float32, device=g.device)
        flat = torch.zeros(1, dtype=torch.int32, device=g.device)
        _lib.check(lib.vr_rdm_finish_f32(g.data_ptr(), n, n, int(d_total), sums[0].data_ptr(), sums[1].data_ptr(), shift.data_ptr(),
                                         float(correction), rdm.data_ptr(), flat.data_ptr(), _lib.stream_ptr()), "vr_rdm_finish_f32")
        return rdm, flat

    return moments, gram, finish


def compute_rdm_ksplit(x_shard: torch.Tensor, d_total: Optional[int] = None, correction: float = 1e-12, group=None,
                       stages=None) -> torch.Tensor:
    """`compute_rdm` (visreps/analysis/rsa.py:59-93) for rows whose FEATURES are split over the ranks of `group`.

    Every rank passes its column shard (n, d_r) of the same n stimuli and receives the full (n, n) RDM:
      1. per-row sum / sum of squares of the shard (fp64)            -> all-reduce of 2n doubles
      2. every rank shifts its shard by the whole-row mean and runs the tensor-core SYRK on it (upper triangle)
      3. the fp32 Gram partials are all-reduced over NVLink BEFORE the non-linear epilogue
      4. covariance / std / clamp / 1 - r / mirror on every rank.
    The n x n Gram is 400 MB at n = 10,000: ~1 ms on NVLink 5, against 91 ms for the single-GPU SYRK of that shape.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if x_shard.ndim > 2:
        x_shard = x_shard.flatten(start_dim=1)
    moments, gram, finish = stages or _rdm_ksplit_stages()
    if stages is None:
        from . import _lib, rsa
        x_shard = rsa._as_cuda_f32_rows(x_shard, _lib.require_cuda())
    sums = moments(x_shard)
    if d_total is None:
        cols = torch.tensor([x_shard.shape[1]], dtype=torch.int64, device=sums.device)
        if world > 1:
            dist.all_reduce(cols, group=group)
        d_total = int(cols.item())
    if world > 1:
        dist.all_reduce(sums, group=group)
    g, shift = gram(x_shard, sums, d_total)
    if world > 1:
        dist.all_reduce(g, group=group)
    rdm, _flat = finish(g, sums, shift, d_total, correction)
    return rdm


def compute_rdm_ksplit_many(shards: Sequence[torch.Tensor], d_totals: Sequence[int], correction: float = 1e-12, group=None,
                            stages=None) -> list:
    """`compute_rdm_ksplit` for several inputs of the same n at once: ONE all-reduce for all the row moments and ONE for all the
    Grams (a collective costs ~50 us of latency whatever its size; eight layers one by one pay it sixteen times)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if not shards:
        return []
    moments, gram, finish = stages or _rdm_ksplit_stages()
    if stages is None:
        from . import _lib, rsa
        dev = _lib.require_cuda()
        shards = [rsa._as_cuda_f32_rows(x.flatten(start_dim=1) if x.ndim > 2 else x, dev) for x in shards]
    n = int(shards[0].shape[0])
    sums = torch.stack([moments(x) for x in shards])                     # (k, 2, n) fp64
    if world > 1:
        dist.all_reduce(sums, group=group)
    grams = torch.empty((len(shards), n, n), dtype=torch.float32, device=sums.device)
    shifts = []
    for k, x in enumerate(shards):
        g, shift = gram(x, sums[k], d_totals[k])
        grams[k] = g
        shifts.append(shift)
    if world > 1:
        dist.all_reduce(grams, group=group)
    return [finish(grams[k], sums[k], shifts[k], d_totals[k], correction)[0] for k in range(len(shards))]


def compute_rdm_sharded(x: torch.Tensor, correction: float = 1e-12, group=None) -> torch.Tensor:
    """Convenience for inputs every rank holds in full: rank r keeps columns column_shard(d, r, world) and the ranks
    cooperate through `compute_rdm_ksplit`."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    x2 = x.flatten(start_dim=1) if x.ndim > 2 else x
    lo, hi = column_shard(int(x2.shape[1]), rank, world)
    return compute_rdm_ksplit(x2[:, lo:hi], d_total=int(x2.shape[1]), correction=correction, group=group)
