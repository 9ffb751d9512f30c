"""Parity of the pair-tile tcgen05 SYRK (csrc/syrk_tc2.cuh: cta_group::2, 256 x 256 tiles), the kernel that builds
every conv-layer RDM at the NSD / THINGS / stress shapes (n > 384, d >= 16384), against the CPU oracle
(reference operation order, rsa.py:59-93), an fp64 truth, and the single-CTA kernel on the same input.

Each case asserts through the library's stage profiler that the pair-tile kernel really ran.
"""
import numpy as np
import pytest
import torch

import visreps_b200 as vb
from visreps_b200 import _lib
from oracle import rsa_oracle as orc

pytestmark = pytest.mark.gpu
RDM_TOL = 1e-5          # north_star: RDM entries within 1e-5 absolute


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


def _post_relu_rows(n: int, d: int, seed: int, dev) -> torch.Tensor:
    """Non-negative, about half zeros, rows sharing latent structure, plus a per-row offset and gain so the
    sampled-shift centring and the tf32 split see rows whose mean is far from zero (SURVEY 8d input recipe)."""
    g = torch.Generator(device=dev).manual_seed(seed)
    latent = torch.randn((n, 16), generator=g, device=dev)
    w = torch.randn((16, d), generator=g, device=dev) / 4.0
    x = torch.relu(latent @ w + torch.randn((n, d), generator=g, device=dev))
    gain = 0.25 + 4.0 * torch.rand((n, 1), generator=g, device=dev)
    offset = 20.0 * torch.rand((n, 1), generator=g, device=dev) * (torch.arange(n, device=dev)[:, None] % 3 == 0)
    return (x * gain + offset).contiguous()


def _truth_f64(x: torch.Tensor) -> torch.Tensor:
    """rdm_truth_f64 of the oracle, evaluated with torch fp64 on the device (the CPU version needs minutes at
    n = 1030 x 193,600); the small case below checks the two agree."""
    z = x.double()
    z = z - z.mean(dim=1, keepdim=True)
    sigma = (z * z).mean(dim=1).add(1e-12).sqrt()
    sigma[sigma < 1e-11] = 1.0
    r = (z @ z.T) / z.shape[1] / (sigma[:, None] * sigma[None, :] + 1e-12)
    r.clamp_(-1, 1).fill_diagonal_(1.0)
    return 1.0 - r


def _rdm_with_families(x, pair: int, waves: int = 1):
    _lib.set_option("syrk_pair", pair)
    _lib.set_option("syrk_waves", waves)
    try:
        _lib.profile_start()
        r = vb.compute_rdm(x)
        torch.cuda.synchronize()
        prof = _lib.profile_stop()
    finally:
        _lib.set_option("syrk_pair", 1)
        _lib.set_option("syrk_waves", 1)
    return r, prof


@pytest.mark.parametrize("n,d,seed", [
    (385, 16384, 11),        # smallest shape that takes pair tiles; n_pad = 512 (127 padded rows)
    (1000, 16423, 12),       # d not a multiple of 32: zero-padded K tail
    (1000, 43264, 13),       # conv4 / conv5 of the NSD bench
    (1030, 64896, 14),       # ragged n (n_pad = 1280), conv3 width
    (1000, 193600, 15),      # conv1 of the NSD bench: 7 K splits of 865 blocks
])
def test_pair_tile_syrk_matches_oracle_fp64_and_single_cta(n, d, seed, dev):
    x = _post_relu_rows(n, d, seed, dev)
    keep = x.clone()
    got, prof = _rdm_with_families(x, pair=1)
    assert prof["rdm_syrk_pair"]["launches"] >= 1 and prof["rdm_syrk"]["launches"] == 0, prof
    assert torch.equal(x, keep), "input mutated"
    assert torch.equal(got, got.T), "RDM must be exactly symmetric"
    assert float(got.diagonal().abs().max()) == 0.0
    assert float(got.min()) >= 0.0 and float(got.max()) <= 2.0
    truth = _truth_f64(x)
    err64 = float((got.double() - truth).abs().max())
    assert err64 <= RDM_TOL, err64
    ref = orc.rdm(x.cpu())                                   # the reference's fp32 CPU order of operations
    err_ref = float((got.cpu() - ref).abs().max())
    assert err_ref <= RDM_TOL, err_ref
    single, prof1 = _rdm_with_families(x, pair=0)
    assert prof1["rdm_syrk"]["launches"] >= 1 and prof1["rdm_syrk_pair"]["launches"] == 0, prof1
    assert float((single - got).abs().max()) <= RDM_TOL
    assert float((single.double() - truth).abs().max()) <= RDM_TOL


def test_pair_tile_syrk_k_split_variants_and_edge_rows(dev):
    """More CTA waves = more (shorter) K splits and partial planes; results stay within the budget of the fp64 truth.
    Rows of zeros, a constant row and a duplicated row exercise the epilogue's std guard through the pair path."""
    n, d = 640, 20000
    x = _post_relu_rows(n, d, 21, dev)
    x[5] = 0.0
    x[17] = 3.5
    x[100] = x[99]
    x[101] = -x[99] + 7.0
    truth = _truth_f64(x)
    outs = []
    for waves in (1, 2, 5):
        got, prof = _rdm_with_families(x, pair=1, waves=waves)
        assert prof["rdm_syrk_pair"]["launches"] >= 1
        assert float((got.double() - truth).abs().max()) <= RDM_TOL
        assert torch.equal(got, got.T)
        outs.append(got)
    assert float((outs[0] - outs[2]).abs().max()) <= RDM_TOL
    got = outs[0]
    assert abs(float(got[99, 100])) <= RDM_TOL and abs(float(got[99, 101]) - 2.0) <= RDM_TOL
    # zero / constant rows: std = sqrt(0 + 1e-12) = 1e-6 escapes the reference's guard, correlations are 0 -> distance 1
    assert torch.all(got[5, torch.arange(n, device=dev) != 5] == 1.0)
    assert torch.all(got[17, torch.arange(n, device=dev) != 17] == 1.0)
    ref = orc.rdm(x.cpu())
    assert float((got.cpu() - ref).abs().max()) <= RDM_TOL


def test_device_fp64_truth_agrees_with_the_oracle_truth(dev):
    x = _post_relu_rows(200, 3000, 31, dev)
    a = _truth_f64(x).cpu().numpy()
    b = orc.rdm_truth_f64(x.cpu())
    assert float(np.abs(a - b).max()) <= 1e-12


def test_packed_upper_triangle_from_the_epilogue(dev):
    """vr_rdm_pearson_triu_f32 / vr_triu_pack_f32: the vector compute_rdm_correlation gathers (rsa.py:111-113), bit for bit."""
    from visreps_b200 import rsa as vrsa
    for n, d in ((1, 5), (2, 9), (67, 130), (500, 20000)):
        x = _post_relu_rows(n, d, 40 + n, dev)
        full = vb.compute_rdm(x)
        iu = torch.triu_indices(n, n, 1, device=dev)
        want = full[iu[0], iu[1]]
        got = vrsa.rdm_upper_triangle(x)
        assert got.shape == want.shape and torch.equal(got, want)
        if n > 1:
            out = torch.empty_like(want)
            _lib.check(_lib.load().vr_triu_pack_f32(full.data_ptr(), n, n, out.data_ptr(), _lib.stream_ptr()), "vr_triu_pack_f32")
            assert torch.equal(out, want)
