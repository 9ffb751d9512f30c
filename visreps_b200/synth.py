"""Synthetic NSD / THINGS-shaped inputs.  No dataset or checkpoint is
available offline, so the benchmark feeds the hot path with random features of the right shapes:
post-ReLU-like model activations and voxel responses that share a latent structure (so RSA scores
are non-trivial)."""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

# torchvision AlexNet (224x224) flattened layer widths, conv1..fc8
ALEXNET_DIMS: Dict[str, int] = {"conv1": 193600, "conv2": 139968, "conv3": 64896, "conv4": 43264,
                                "conv5": 43264, "fc6": 4096, "fc7": 4096, "fc8": 1000}
# plausible NSD ROI voxel counts (assumption stated in SURVEY.md section 8d)
NSD_ROIS: Dict[str, int] = {"V1": 1400, "V2": 1300, "V3": 1100, "hV4": 600, "FFA": 500, "PPA": 400,
                            "early": 5900, "ventral": 7600}
LATENT = 64


def _gen(device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def latent(n: int, device, seed: int = 1) -> torch.Tensor:
    return torch.randn(n, LATENT, device=device, generator=_gen(device, seed))


def layer_features(z: torch.Tensor, d: int, seed: int, chunk: int = 32768) -> torch.Tensor:
    """relu(Z W + eps), W ~ N(0,1)/8, generated in column chunks: non-negative, about half zeros."""
    dev = z.device
    g = _gen(dev, seed)
    out = torch.empty(z.shape[0], d, device=dev, dtype=torch.float32)
    for c0 in range(0, d, chunk):
        c1 = min(d, c0 + chunk)
        w = torch.randn(LATENT, c1 - c0, device=dev, generator=g) / 8
        out[:, c0:c1] = torch.relu(z @ w + torch.randn(z.shape[0], c1 - c0, device=dev, generator=g))
    return out


def roi_responses(z: torch.Tensor, v: int, seed: int) -> torch.Tensor:
    dev = z.device
    g = _gen(dev, seed)
    a = torch.randn(LATENT, v, device=dev, generator=g)
    return z @ a + 0.5 * torch.randn(z.shape[0], v, device=dev, generator=g) * (LATENT ** 0.5)


def nsd_workload(n: int = 1000, device="cpu", seed: int = 0, layers=None, rois=None
                 ) -> Tuple[List[Tuple[str, torch.Tensor]], List[Tuple[str, torch.Tensor]]]:
    """[(layer, (n, d) features)], [(roi, (n, V) responses)] for one model."""
    device = torch.device(device)
    z = latent(n, device, seed * 1000 + 1)
    layers = ALEXNET_DIMS if layers is None else layers
    rois = NSD_ROIS if rois is None else rois
    feats = [(k, layer_features(z, d, seed * 1000 + 10 + i)) for i, (k, d) in enumerate(layers.items())]
    # ROI data do not depend on the model seed: every model is compared with the same brain
    zb = latent(n, device, 1)
    mix = 0.7 * z + 0.3 * zb if seed else z
    neur = [(k, roi_responses(mix, v, 500 + i)) for i, (k, v) in enumerate(rois.items())]
    return feats, neur
