"""Device-resident hand-off from the activation producer.

The reference's producer runs the network on the GPU and then pushes every activation through `.cpu()`
(visreps/models/utils.py:344 in `get_activations`, :391 in `extract_single_layer`); with the RSA numerics on the
device that round trip is pure PCIe cost (2.05 GB per NSD-shape model, 37 of the 41 ms of an end-to-end step).
The two functions below have the reference's signatures and return the same (tensor(s), ids) -- fp32, same row
order -- but the tensors stay on `device`.  `visreps_b200.install(device_producer=True)` binds them over
`visreps.models.utils.{get_activations, extract_single_layer}`; everything downstream of them on the RSA path
(`_align_stimulus_level`, `compute_rdm`, `_eval_rsa`, `compute_rsa`, `_concept_average_exact`) accepts device tensors.

The network, its forward hooks and the sparse random projection are the reference's own (PyTorch; not optimisation
scope): `srp_factory` is the reference's `get_srp_transformer`, looked up on the host module when not given.
"""
from __future__ import annotations

import sys
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

SRP_K = 4096                                   # models/utils.py:305
SRP_CACHE = "model_checkpoints/srp_cache"      # models/utils.py:305


def _flat2d(out: torch.Tensor) -> torch.Tensor:
    return out.reshape(out.size(0), -1)


def _srp_factory(explicit: Optional[Callable]) -> Callable:
    if explicit is not None:
        return explicit
    for name in ("visreps.models.utils", "visreps.analysis.sparse_random_projection"):
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "get_srp_transformer"):
            return mod.get_srp_transformer
    raise RuntimeError("get_srp_transformer not found: import visreps.models.utils first or pass srp_factory=")


def get_activations(model, dataloader, device, srp_factory: Optional[Callable] = None
                    ) -> Tuple[Dict[str, torch.Tensor], List]:
    """`get_activations` (models/utils.py:281-347) without the host round trip: SRP-projected activations of every
    tapped layer, (n, min(4096, D)) fp32 per layer, rows in dataloader order, resident on `device`."""
    model.eval()
    try:
        probe_imgs, _ = next(iter(dataloader))
    except StopIteration:
        return {}, []
    with torch.no_grad():
        probe_out = model(probe_imgs.to(device))
    factory = _srp_factory(srp_factory)
    proj: Dict[str, torch.Tensor] = {}
    for name, out in probe_out.items():
        D = _flat2d(out).size(1)
        coo = factory(D=D, k=min(SRP_K, D), density=None, seed=None, cache_dir=SRP_CACHE).components_.tocoo()
        idx = torch.from_numpy(np.vstack([coo.row, coo.col])).long()
        proj[name] = torch.sparse_coo_tensor(idx, torch.from_numpy(coo.data).float(), coo.shape).to(device)
    total = len(dataloader.dataset) if hasattr(dataloader, "dataset") else None
    store: Dict[str, torch.Tensor] = {}
    chunks: Dict[str, list] = {}
    ids: List = []
    filled = 0
    with torch.no_grad():
        for imgs, keys in dataloader:
            ids.extend(keys)
            feats = model(imgs.to(device, non_blocking=True))
            b = imgs.size(0)
            for name, out in feats.items():
                p = proj.get(name)
                if p is not None:
                    out = torch.sparse.mm(p, _flat2d(out).float().t()).t()
                out = out.float()
                if total is None:
                    chunks.setdefault(name, []).append(out)
                    continue
                if name not in store:                      # one device buffer per layer, filled batch by batch
                    store[name] = torch.empty((total,) + tuple(out.shape[1:]), dtype=torch.float32, device=device)
                store[name][filled:filled + b] = out
            filled += b
    if total is None:
        return {n: torch.cat(c, 0) for n, c in chunks.items()}, ids
    return {n: t[:filled] for n, t in store.items()}, ids


def extract_single_layer(model, dataloader, device, layer_name: str, stimulus_ids: Optional[List[str]] = None
                         ) -> Tuple[torch.Tensor, List]:
    """`extract_single_layer` (models/utils.py:350-404) without the host round trip: exact (un-projected) activations
    of one layer, flattened, fp32 on `device`; rows follow `stimulus_ids` when given."""
    model.eval()
    parts, all_ids = [], []
    with torch.no_grad():
        for imgs, keys in dataloader:
            all_ids.extend(keys)
            out = model(imgs.to(device, non_blocking=True))[layer_name]
            parts.append(_flat2d(out).float())
    acts = torch.cat(parts, 0) if parts else torch.empty((0, 0), dtype=torch.float32, device=device)
    if stimulus_ids is not None:
        where = {str(k): i for i, k in enumerate(all_ids)}
        keep = [where[str(s)] for s in stimulus_ids if str(s) in where]
        acts = acts[torch.as_tensor(keep, dtype=torch.long, device=acts.device)]
        all_ids = [all_ids[i] for i in keep]
    return acts, all_ids
