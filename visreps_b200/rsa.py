"""Host-side mirror of the reference RSA interface, backed by the sm_100a CUDA library.

Same names, argument meaning, return types and error behaviour as
/root/reference/visreps/analysis/rsa.py (compute_rdm :59-93, compute_rdm_correlation :96-129,
compute_rsa :132-281, _rank :50-52, _concept_average_exact :284-305) plus the bootstrap helper
the reference inlines three times (visreps/evals.py:355-373, :506-522) and spells out once
(experiments/reconstruction_analysis/run_reconstruction.py:137-158).

Division of labour: this module validates arguments, draws the bootstrap index sets on the host
from numpy's RandomState exactly as the reference does, moves tensors, and calls the C ABI
(include/visreps_b200.h).  All arithmetic on the path runs in the CUDA library.  There is no CPU
fallback: without the library or without a CUDA device every function raises.
"""
from __future__ import annotations

import collections
import ctypes as C
import logging
from typing import Dict, List, Optional, Sequence, Tuple, TYPE_CHECKING

import numpy as np
import torch

from . import _lib

if TYPE_CHECKING:  # pragma: no cover
    from .alignment import AlignmentData

logger = logging.getLogger(__name__)

SUBSAMPLE_FRACTION = 0.9  # rsa.py:234 / evals.py:358
_METHODS = ("pearson", "spearman", "kendall")
_MAX_GROUP = 32  # vectors per side in one grouped library call (library limit: 64 in total)


# --------------------------------------------------------------------------- helpers
def _as_cuda_f32_rows(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    """fp32 CUDA view/copy whose rows are unit-stride (row pitch may exceed the width)."""
    x = t.detach()
    if x.device != dev:
        x = x.to(dev, non_blocking=True)
    if x.dtype != torch.float32:
        x = x.float()
    if x.dim() != 2:
        raise ValueError("expected a 2-D tensor")
    if x.size(1) > 1 and x.stride(1) != 1 or x.stride(0) < x.size(1):
        x = x.contiguous()
    return x


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class PresortedRDM:
    """An RDM's strict upper triangle sorted once on the device (opaque library blob).

    Every comparison / bootstrap replicate of this RDM is ranked from the blob without
    sorting again.  Keeps the source RDM alive for the Pearson path.
    """

    def __init__(self, rdm: torch.Tensor, _defer: bool = False):
        lib = _lib.load()
        dev = _lib.require_cuda()
        self.rdm = _as_cuda_f32_rows(rdm, dev)
        if self.rdm.size(0) != self.rdm.size(1):
            raise ValueError("RDM must be square")
        self.n = int(self.rdm.size(0))
        self.blob = torch.empty(lib.vr_presort_bytes(self.n), dtype=torch.uint8, device=dev)
        if not _defer:
            PresortedRDM._sort_batch([self])

    @classmethod
    def from_blob(cls, blob: torch.Tensor, n: int, rdm: Optional[torch.Tensor] = None) -> "PresortedRDM":
        """Adopt a blob presorted elsewhere (another rank): `vr_presort_bytes(n)` bytes on this device."""
        lib = _lib.load()
        if blob.dtype != torch.uint8 or blob.numel() != lib.vr_presort_bytes(n) or not blob.is_cuda or not blob.is_contiguous():
            raise ValueError("blob must be a contiguous CUDA uint8 tensor of vr_presort_bytes(n) bytes")
        self = cls.__new__(cls)
        self.rdm, self.n, self.blob = rdm, int(n), blob
        return self

    @staticmethod
    def _sort_batch(items: "Sequence[PresortedRDM]") -> None:
        lib = _lib.load()
        dev = items[0].blob.device
        n, ld = items[0].n, items[0].rdm.stride(0)
        for k0 in range(0, len(items), 64):
            part = items[k0:k0 + 64]
            for it in part:
                if it.rdm.stride(0) != ld:
                    it.rdm = it.rdm.contiguous()
            if any(it.rdm.stride(0) != ld for it in part):
                for it in part:
                    it.rdm = it.rdm.contiguous()
                ld = n
            ws = _lib.workspace(lib.vr_presort_batch_workspace_bytes(n, len(part)), dev)
            _lib.check(lib.vr_presort_triu_batch_f32(_ptr_array([it.rdm for it in part]), len(part), n, ld,
                                                     _ptr_array([it.blob for it in part]), ws.data_ptr(), ws.numel(),
                                                     _lib.stream_ptr()), "vr_presort_triu_batch_f32")

    @classmethod
    def batch(cls, rdms: Sequence[torch.Tensor]) -> "List[PresortedRDM]":
        """Presort several RDMs of the same size with one batched device sort."""
        items = [cls(r, _defer=True) for r in rdms]
        if items:
            if any(it.n != items[0].n for it in items):
                raise ValueError("RDMs must share the same 2-D shape")
            cls._sort_batch(items)
        return items


def full_mask(n: int) -> np.ndarray:
    """All-stimuli membership mask (uint32 words): gives the point estimate."""
    w = (n + 31) // 32
    m = np.full(w, 0xFFFFFFFF, dtype=np.uint32)
    if n % 32:
        m[-1] = (1 << (n % 32)) - 1
    return m


def build_masks(index_sets: Optional[np.ndarray], n: int, device: torch.device, with_full: bool = True) -> torch.Tensor:
    """(R, W) int32 device tensor of membership bit masks; row 0 is the full mask if with_full.

    index_sets: host (B, m) integer array, the reference's `rng.choice(n, m, replace=False)` draws.
    The index array is staged in pinned host memory and read by the mask kernel directly over PCIe
    (unified addressing), so it never queues behind bulk host->device copies on the DMA engine.
    """
    lib = _lib.load()
    w = lib.vr_mask_words(n)
    b = 0 if index_sets is None else len(index_sets)
    cache_key = (str(device), bool(with_full))
    if isinstance(index_sets, HostMasks) and index_sets.cached and cache_key in index_sets.device_copies:
        return index_sets.device_copies[cache_key]                   # read-only by contract: every consumer only reads masks
    masks = torch.empty((b + (1 if with_full else 0), w), dtype=torch.int32, device=device)
    if with_full:
        masks[0].fill_(-1)
        if n % 32:
            masks[0, w - 1:].fill_((1 << (n % 32)) - 1)
    if b and isinstance(index_sets, HostMasks):
        if index_sets.n != n:
            raise ValueError("masks were drawn for a different number of stimuli")
        sub = masks[1:] if with_full else masks
        _lib.check(lib.vr_fetch_masks(index_sets.words.data_ptr(), b, n, sub.data_ptr(), _lib.stream_ptr()), "vr_fetch_masks")
        _keep_alive(index_sets.words)
        if index_sets.cached:
            index_sets.device_copies[cache_key] = masks
    elif b:
        idx_np = np.ascontiguousarray(index_sets, dtype=np.int32)
        if idx_np.size and (int(idx_np.min()) < 0 or int(idx_np.max()) >= n):
            raise ValueError("bootstrap index out of range")
        idx = torch.from_numpy(idx_np).pin_memory()
        sub = masks[1:] if with_full else masks
        _lib.check(lib.vr_build_masks(idx.data_ptr(), b, int(idx.shape[1]), n, sub.data_ptr(), _lib.stream_ptr()),
                   "vr_build_masks")
        _keep_alive(idx)
    return masks


_pinned_in_flight: list = []


def _keep_alive(t: torch.Tensor) -> None:
    """Hold a pinned staging buffer until the kernels reading it have run (released at the next sync point)."""
    _pinned_in_flight.append((t, torch.cuda.Event()))
    _pinned_in_flight[-1][1].record()
    while len(_pinned_in_flight) > 1 and _pinned_in_flight[0][1].query():
        _pinned_in_flight.pop(0)


def draw_index_sets(n: int, n_bootstrap: int, seed: int = 42, rng: Optional[np.random.RandomState] = None) -> np.ndarray:
    """Host RNG stream of the reference bootstrap: B x int(0.9 n) draws without replacement,
    in draw order (evals.py:356-364; rsa.py:249-251).

    The B `rng.choice(n, m, replace=False)` calls are replayed by the library's host routine on the
    generator's own MT19937 state (bit-identical, ~20x faster than the Python loop); `rng` is left in
    exactly the state the reference's loop would leave it in."""
    rng = np.random.RandomState(seed) if rng is None else rng
    m = int(n * SUBSAMPLE_FRACTION)
    out = np.empty((n_bootstrap, m), dtype=np.int32)
    if n_bootstrap == 0:
        return out
    state = rng.get_state()
    if state[0] != "MT19937" or n < 1:
        for b in range(n_bootstrap):
            out[b] = rng.choice(n, size=m, replace=False)
        return out
    key = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
    pos = C.c_int(int(state[2]))
    _lib.check(_lib.load().vr_host_draw_subsamples(key.ctypes.data, C.byref(pos), n, m, n_bootstrap, out.ctypes.data),
               "vr_host_draw_subsamples")
    rng.set_state((state[0], key, pos.value, state[3], state[4]))
    return out


def percentile_ci(scores: np.ndarray, lo: float = 2.5, hi: float = 97.5) -> Tuple[np.ndarray, np.ndarray]:
    """`np.percentile(scores, [lo, hi], axis=-1)` (default 'linear' method) bit for bit -- the CI of
    evals.py:371-372 / rsa.py:262-263 -- from one vectorised sort of the last axis instead of NumPy's
    generic quantile machinery (which costs ~4 ms for 64 x 1000 scores, on the critical path of a step)."""
    scores = np.asarray(scores, dtype=np.float64)
    b = scores.shape[-1]
    if b == 0:
        nan = np.full(scores.shape[:-1], np.nan)
        return nan, nan.copy()
    srt = np.sort(scores, axis=-1)                      # NaNs last; a row with any NaN yields NaN, as NumPy does
    bad = np.isnan(srt[..., -1])
    out = []
    for q in (lo, hi):
        virt = q / 100.0 * (b - 1)
        k = int(np.floor(virt))
        t = virt - k
        lower, upper = srt[..., k], srt[..., min(k + 1, b - 1)]
        diff = upper - lower
        val = upper - diff * (1 - t) if t >= 0.5 else lower + diff * t          # numpy's _lerp
        out.append(np.where(bad, np.nan, val))
    return out[0], out[1]


def percentile_ci_device(scores: torch.Tensor, lo: float = 2.5, hi: float = 97.5) -> Optional[torch.Tensor]:
    """`percentile_ci` on the device (vr_percentile_ci_f64: same bits as np.percentile): scores is a (..., B) fp64 CUDA tensor
    whose last axis is contiguous and whose leading axes are evenly strided (a slice [..., 1:] of a contiguous tensor is fine).
    Returns a (2, ...) tensor [low, high], or None when B is outside what the kernel sorts in shared memory (caller falls back
    to the host routine on the copied scores)."""
    lib = _lib.load()
    b = int(scores.shape[-1])
    lead = tuple(scores.shape[:-1])
    rows = int(np.prod(lead)) if lead else 1
    if b < 1 or b > lib.vr_percentile_ci_max_replicates() or scores.dtype != torch.float64 or not scores.is_cuda:
        return None
    if scores.stride(-1) != 1:
        return None
    row_stride = int(scores.stride(-2)) if scores.ndim >= 2 else b
    expect = row_stride
    for d in range(scores.ndim - 2, -1, -1):                      # leading axes must collapse to one stride
        if scores.shape[d] != 1 and int(scores.stride(d)) != expect:
            return None
        expect *= int(scores.shape[d])
    out = torch.empty((2,) + lead, dtype=torch.float64, device=scores.device)
    if rows:
        _lib.check(lib.vr_percentile_ci_f64(scores.data_ptr(), rows, b, row_stride, float(lo), float(hi),
                                            out[0].data_ptr(), out[1].data_ptr(), _lib.stream_ptr()), "vr_percentile_ci_f64")
    return out


class HostMasks:
    """Page-locked (B, W) uint32 membership masks drawn on the host (row b = kept stimuli of draw b)."""

    def __init__(self, words: torch.Tensor, n: int):
        self.words, self.n = words, n
        self.device_copies: Dict[Tuple[str, bool], torch.Tensor] = {}      # filled by build_masks for cached draws only
        self.cached = False

    def __len__(self) -> int:
        return int(self.words.shape[0])


# The draws are a pure function of (n, B, seed) when the caller passes no generator of its own: the secondary callers
# (run_reconstruction.py:137-158: 15 pca_k x pairs, one bootstrap_rdm_correlation each, same seed) replay the same 1000
# draws for every pair, ~2 ms of host time per call against 0.2-3 ms of device work.  A few of them are kept (pinned host
# words + their device copy).  MASK_CACHE = False replays every call (bench.py does, so that no step skips the draw).
MASK_CACHE = True
_MASK_CACHE_MAX = 8
_mask_cache: "collections.OrderedDict[tuple, HostMasks]" = collections.OrderedDict()


def draw_subsample_masks(n: int, n_bootstrap: int, seed: int = 42, rng: Optional[np.random.RandomState] = None,
                         replicate_range: Optional[Tuple[int, int]] = None) -> HostMasks:
    """The index draws of `draw_index_sets`, delivered as membership masks only (what the device needs):
    same generator stream, same final generator state, but each shuffle is carried out only for the
    int(0.1 n)+ slots that end up dropped while the rest of its draws are merely consumed, so the host
    is off the critical path ~3x sooner.  `replicate_range` keeps only that slice of the draws (a rank's
    share, visreps_b200/dist.py) -- the generator still advances through all of them."""
    b0, b1 = (0, n_bootstrap) if replicate_range is None else (int(replicate_range[0]), int(replicate_range[1]))
    b0, b1 = max(0, min(b0, n_bootstrap)), max(0, min(b1, n_bootstrap))
    b1 = max(b0, b1)
    key = (int(n), int(n_bootstrap), int(seed), b0, b1, float(SUBSAMPLE_FRACTION)) if (rng is None and MASK_CACHE) else None
    if key is not None and key in _mask_cache:
        _mask_cache.move_to_end(key)
        return _mask_cache[key]
    hm = _draw_subsample_masks(n, n_bootstrap, np.random.RandomState(seed) if rng is None else rng, b0, b1)
    if key is not None:
        hm.cached = True
        _mask_cache[key] = hm
        while len(_mask_cache) > _MASK_CACHE_MAX:
            _mask_cache.popitem(last=False)
    return hm


def _draw_subsample_masks(n: int, n_bootstrap: int, rng: np.random.RandomState, b0: int, b1: int) -> HostMasks:
    m = int(n * SUBSAMPLE_FRACTION)
    lib = _lib.load()
    w = lib.vr_mask_words(n)
    words = torch.empty((b1 - b0, w), dtype=torch.int32, pin_memory=torch.cuda.is_available())
    if n_bootstrap == 0:
        return HostMasks(words, n)
    state = rng.get_state()
    if state[0] != "MT19937" or n < 1:
        idx = np.stack([rng.choice(n, size=m, replace=False) for _ in range(n_bootstrap)])[b0:b1]
        bits = np.zeros((b1 - b0, w * 32), dtype=bool)
        np.put_along_axis(bits, idx.astype(np.int64), True, axis=1)
        words.copy_(torch.from_numpy(np.packbits(bits, axis=1, bitorder="little").view(np.int32).reshape(b1 - b0, w)))
        return HostMasks(words, n)
    key = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
    pos = C.c_int(int(state[2]))
    _lib.check(lib.vr_host_draw_subsample_masks(key.ctypes.data, C.byref(pos), n, m, n_bootstrap, b0, b1, words.data_ptr()),
               "vr_host_draw_subsample_masks")
    rng.set_state((state[0], key, pos.value, state[3], state[4]))
    return HostMasks(words, n)


def _ptr_array(tensors: Sequence[torch.Tensor]):
    arr = (C.c_void_p * len(tensors))()
    for k, t in enumerate(tensors):
        arr[k] = t.data_ptr()
    return arr


def _int_array(values: Sequence[int]):
    arr = (C.c_int * len(values))()
    for k, v in enumerate(values):
        arr[k] = int(v)
    return arr


import os as _os

# "sweep": one warp per (canonical pair, 32 replicates), csrc/sweep.cu (default).  "legacy": one CTA per (RDM, replicate) with
# the count table in shared memory, csrc/rank.cu -- kept for A/B measurements (VR_SPEARMAN_ENGINE=legacy).
SPEARMAN_ENGINE = _os.environ.get("VR_SPEARMAN_ENGINE", "sweep")


# Host inputs at a size whose count table fits shared memory (n <= 1277) stay on the legacy kernels: what is exposed after
# the last host->device copy is then one matrix ranked against the materialised side (3 ms at the NSD shape, HBM streaming),
# where a sweep of one matrix against eight costs 4 ms of table-row traffic.  Everything else takes the sweep.
_LEGACY_SMEM_TABLE_MAX_N = 1277


class SpearmanJob:
    """Count tables of up to `slots` presorted RDMs for one set of masks (library-opaque device buffer), and the grouped
    Spearman sweep over them.  RDMs can be added as they become available; every (A, B) pair of a sweep shares the lookups."""

    def __init__(self, masks: torch.Tensor, n: int, slots: int):
        lib = _lib.load()
        self.n, self.R, self.slots, self.masks = int(n), int(masks.shape[0]), int(slots), masks
        self.buf = torch.empty(lib.vr_spearman_job_bytes(self.n, self.slots, self.R), dtype=torch.uint8, device=masks.device)
        self.items: List[Optional[PresortedRDM]] = [None] * self.slots
        _lib.check(lib.vr_spearman_job_init(masks.data_ptr(), self.n, self.R, self.slots, self.buf.data_ptr(), self.buf.numel(),
                                            _lib.stream_ptr()), "vr_spearman_job_init")

    @staticmethod
    def bytes_needed(n: int, slots: int, R: int) -> int:
        return int(_lib.load().vr_spearman_job_bytes(int(n), int(slots), int(R)))

    def add(self, items: Sequence[PresortedRDM], slot0: int) -> None:
        lib = _lib.load()
        if any(p.n != self.n for p in items):
            raise ValueError("RDMs must share the same 2-D shape")
        _lib.check(lib.vr_spearman_job_add(self.buf.data_ptr(), self.n, self.R, self.slots, _ptr_array([p.blob for p in items]),
                                           len(items), int(slot0), _lib.stream_ptr()), "vr_spearman_job_add")
        for k, p in enumerate(items):
            self.items[slot0 + k] = p

    def sweep(self, slots_a: Sequence[int], slots_b: Sequence[int], out: torch.Tensor, offset: int, stride_a: int, stride_b: int) -> None:
        """rho of every (a, b, replicate) into the flat fp64 tensor `out` at offset + ia * stride_a + ib * stride_b + r."""
        lib = _lib.load()
        dev = self.buf.device
        ws = _lib.workspace(lib.vr_spearman_sweep_workspace_bytes(self.n, len(slots_a), len(slots_b), self.R), dev, "sweep")
        _lib.check(lib.vr_spearman_job_sweep(self.buf.data_ptr(), self.n, self.R, self.slots,
                                             _ptr_array([self.items[k].blob for k in slots_a]), _int_array(slots_a), len(slots_a),
                                             _ptr_array([self.items[k].blob for k in slots_b]), _int_array(slots_b), len(slots_b),
                                             out.data_ptr() + 8 * int(offset), int(stride_a), int(stride_b), ws.data_ptr(), ws.numel(),
                                             _lib.stream_ptr()), "vr_spearman_job_sweep")


def _spearman_sweep_grouped(model: Sequence[PresortedRDM], neural: Sequence[PresortedRDM], masks: torch.Tensor, out: torch.Tensor) -> None:
    """out (nA, nB, R) fp64 <- Spearman of every model x neural x mask, replicates chunked to the memory at hand."""
    n = model[0].n
    nA, nB, R = len(model), len(neural), int(masks.shape[0])
    dev = masks.device
    key = ("chunk", n, nA, nB, R, dev.index)
    chunk = _fit_cache.get(key)
    if chunk is None:       # cudaMemGetInfo costs 0.2-1.5 ms: asked once per problem shape, not once per call
        free, _total = torch.cuda.mem_get_info(dev)
        cached = torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
        budget = max(int(0.5 * (free + cached)), 1 << 28)
        per_batch = max(1, SpearmanJob.bytes_needed(n, nA + nB, 64) - SpearmanJob.bytes_needed(n, nA + nB, 32))
        chunk = _fit_cache[key] = max(32, min((R + 31) // 32 * 32, budget // per_batch * 32))
    flat = out.view(-1)
    for r0 in range(0, R, chunk):
        rc = min(chunk, R - r0)
        job = SpearmanJob(masks[r0:r0 + rc], n, nA + nB)
        job.add(list(model) + list(neural), 0)
        if rc == R:
            job.sweep(range(nA), range(nA, nA + nB), flat, 0, nB * R, R)
        else:
            part = torch.empty((nA, nB, rc), dtype=torch.float64, device=dev)
            job.sweep(range(nA), range(nA, nA + nB), part.view(-1), 0, nB * rc, rc)
            out[:, :, r0:r0 + rc] = part


def compare_grouped(model: Sequence[PresortedRDM], neural: Sequence[PresortedRDM], method: str,
                    masks: torch.Tensor, kept_stimuli: Optional[int] = None) -> torch.Tensor:
    """(nA, nB, R) fp64 device tensor: every model RDM x every neural RDM x every mask.

    kept_stimuli: how many stimuli a subsample mask keeps, when the caller knows (the bootstrap: int(0.9 n)).  Kendall jobs of
    192 or more masks build the merge tree of each RDM pair once and count every replicate without sorting (csrc/kendall.cu,
    tree mode); smaller jobs sort, counting ties and inversions among the DROPPED entries when most are kept (complement mode)
    and among the kept ones otherwise; the answer is the same bits every way."""
    lib = _lib.load()
    method = method.lower()
    if method not in _METHODS:
        raise ValueError("correlation must be 'Pearson', 'Spearman', or 'Kendall'")
    n = model[0].n
    if any(p.n != n for p in list(model) + list(neural)):
        raise ValueError("RDMs must share the same 2-D shape")
    dev = masks.device
    nA, nB, R = len(model), len(neural), int(masks.shape[0])
    out = torch.empty((nA, nB, R), dtype=torch.float64, device=dev)
    for a0 in range(0, nA, _MAX_GROUP):
        for b0 in range(0, nB, _MAX_GROUP):
            ma, mb = model[a0:a0 + _MAX_GROUP], neural[b0:b0 + _MAX_GROUP]
            full = len(ma) == nA and len(mb) == nB
            part = out if full else torch.empty((len(ma), len(mb), R), dtype=torch.float64, device=dev)
            if method == "kendall":
                _lib.set_option("kendall_complement", 1 if kept_stimuli is None or kept_stimuli * kept_stimuli >= 0.5 * n * n else 0)
                ws = _lib.workspace(lib.vr_kendall_batch_workspace_bytes(n, R), dev)
                for ia, a in enumerate(ma):
                    for ib, b in enumerate(mb):
                        ra = a.rdm if a.rdm.stride(0) == n else a.rdm.contiguous()
                        rb = b.rdm if b.rdm.stride(0) == n else b.rdm.contiguous()
                        _lib.check(lib.vr_kendall_tau_a(ra.data_ptr(), rb.data_ptr(), n, n, masks.data_ptr(), R,
                                                        part[ia, ib].data_ptr(), ws.data_ptr(), ws.numel(),
                                                        _lib.stream_ptr()), "vr_kendall_tau_a")
            elif method == "spearman" and SPEARMAN_ENGINE == "sweep":
                _spearman_sweep_grouped(ma, mb, masks, part)
            elif method == "spearman":
                pa, pb = _ptr_array([p.blob for p in ma]), _ptr_array([p.blob for p in mb])
                ws = _lib.workspace(lib.vr_bootstrap_spearman_workspace_bytes(n, len(ma), len(mb), R), dev)
                _lib.check(lib.vr_bootstrap_spearman(pa, len(ma), pb, len(mb), n, masks.data_ptr(), R, part.data_ptr(),
                                                     ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "vr_bootstrap_spearman")
            else:
                ld = ma[0].rdm.stride(0)
                rows = [p.rdm if p.rdm.stride(0) == ld else p.rdm.contiguous() for p in list(ma) + list(mb)]
                if any(r.stride(0) != ld for r in rows):
                    rows = [r.contiguous() for r in rows]
                    ld = n
                pa, pb = _ptr_array(rows[:len(ma)]), _ptr_array(rows[len(ma):])
                ws = _lib.workspace(lib.vr_bootstrap_pearson_workspace_bytes(n, len(ma), len(mb), R), dev)
                for r0 in range(0, R, 65535):
                    rc = min(65535, R - r0)
                    sub = part if rc == R else torch.empty((len(ma), len(mb), rc), dtype=torch.float64, device=dev)
                    _lib.check(lib.vr_bootstrap_pearson(pa, len(ma), pb, len(mb), n, ld, masks[r0:r0 + rc].data_ptr(), rc,
                                                        sub.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "vr_bootstrap_pearson")
                    if sub is not part:
                        part[:, :, r0:r0 + rc] = sub
            if not full:
                out[a0:a0 + len(ma), b0:b0 + len(mb)] = part
    return out


# --------------------------------------------------------------------------- public API
def _rank(x: torch.Tensor) -> torch.Tensor:
    """Row-wise 0-based ordinal ranks, fp32 (reference: rsa.py:50-52)."""
    lib = _lib.load()
    dev = _lib.require_cuda()
    src = x.device
    xs = _as_cuda_f32_rows(x, dev)
    n, d = int(xs.size(0)), int(xs.size(1))
    out = torch.empty((n, d), dtype=torch.float32, device=dev)
    if n and d:
        ws = _lib.workspace(lib.vr_rank_rows_workspace_bytes(n, d), dev)
        _lib.check(lib.vr_rank_rows_ordinal_f32(xs.data_ptr(), n, d, xs.stride(0), out.data_ptr(), ws.data_ptr(),
                                                ws.numel(), _lib.stream_ptr()), "vr_rank_rows_ordinal_f32")
    return out.to(src)


def compute_rdm(representations: torch.Tensor, *, correlation: str = "Pearson", correction: float = 1e-12) -> torch.Tensor:
    """(n x n) fp32 RDM = 1 - correlation between rows (reference: rsa.py:59-93).

    Diagonal exactly 0, off-diagonal in [0, 2], exactly symmetric.  The input is not modified.
    The result lives on the input's device.
    """
    kind = correlation.lower()
    if kind not in ("pearson", "spearman"):
        raise ValueError("correlation must be 'Pearson' or 'Spearman'")
    lib = _lib.load()
    dev = _lib.require_cuda()
    src = representations.device
    x = _as_cuda_f32_rows(representations, dev)
    if kind == "spearman":
        x = _rank(x)
    n, d = int(x.size(0)), int(x.size(1))
    rdm = torch.empty((n, n), dtype=torch.float32, device=dev)
    if n == 0:
        return rdm.to(src)
    if d == 0:
        raise ValueError("representations need at least one feature")
    flat = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = _lib.workspace(lib.vr_rdm_workspace_bytes(n, d), dev)
    _lib.check(lib.vr_rdm_pearson_f32(x.data_ptr(), n, d, x.stride(0), float(correction), rdm.data_ptr(), flat.data_ptr(),
                                      ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "vr_rdm_pearson_f32")
    if logger.isEnabledFor(logging.WARNING):
        nz = int(flat.item())
        if nz:
            logger.warning("%d/%d rows have ~zero variance for %s RDM", nz, n, correlation)
    return rdm.to(src)


def rdm_upper_triangle(representations: torch.Tensor, correction: float = 1e-12) -> torch.Tensor:
    """The strict upper triangle of `compute_rdm(representations)` in `torch.triu_indices(n, n, 1)` order (rsa.py:111-113),
    written by the SYRK epilogue itself: the (n, n) matrix is never formed."""
    lib = _lib.load()
    dev = _lib.require_cuda()
    src = representations.device
    x = _as_cuda_f32_rows(_flat(representations), dev)
    n, d = int(x.size(0)), int(x.size(1))
    out = torch.empty(n * (n - 1) // 2 if n > 1 else 0, dtype=torch.float32, device=dev)
    if n > 1:
        ws = _lib.workspace(lib.vr_rdm_workspace_bytes(n, d), dev)
        _lib.check(lib.vr_rdm_pearson_triu_f32(x.data_ptr(), n, d, x.stride(0), float(correction), None, out.data_ptr(), None,
                                               ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "vr_rdm_pearson_triu_f32")
    return out.to(src)


def compute_rdm_correlation(rdm1: torch.Tensor, rdm2: torch.Tensor, *, correlation: str = "Kendall") -> float:
    """Correlation of two RDMs over the strict upper triangle (reference: rsa.py:96-129).

    NaN (not an exception) when the correlation is undefined."""
    if rdm1.shape != rdm2.shape or rdm1.ndim != 2:
        raise ValueError("RDMs must share the same 2-D shape")
    n = rdm1.size(0)
    if n <= 1:
        logger.warning("RDM dimension <= 1; correlation undefined")
        return float("nan")
    kind = correlation.lower()
    if kind not in _METHODS:
        raise ValueError("correlation must be 'Pearson', 'Spearman', or 'Kendall'")
    dev = _lib.require_cuda()
    a, b = PresortedRDM(rdm1) if kind == "spearman" else _RawRDM(rdm1), \
        PresortedRDM(rdm2) if kind == "spearman" else _RawRDM(rdm2)
    masks = build_masks(None, int(n), dev, with_full=True)
    val = float(compare_grouped([a], [b], kind, masks)[0, 0, 0].item())
    if np.isnan(val):
        logger.warning("NaN returned for %s correlation", correlation)
        return float("nan")
    return val


_COMPARE_IMPL = compute_rdm_correlation


class _RawRDM:
    """RDM on the device without a presort (Pearson comparisons need only the matrix)."""

    def __init__(self, rdm: torch.Tensor):
        dev = _lib.require_cuda()
        self.rdm = _as_cuda_f32_rows(rdm, dev)
        if self.rdm.size(0) != self.rdm.size(1):
            raise ValueError("RDM must be square")
        self.n = int(self.rdm.size(0))


def _prepare(rdm, method: str):
    if isinstance(rdm, (PresortedRDM, _RawRDM)):
        return rdm
    return PresortedRDM(rdm) if method == "spearman" else _RawRDM(rdm)


def _prepare_many(rdms: Sequence, method: str) -> list:
    """Presort (Spearman) or wrap (Pearson) a list of RDMs; raw tensors are presorted in one batch."""
    if method != "spearman":
        return [_prepare(r, method) for r in rdms]
    out = [r if isinstance(r, PresortedRDM) else None for r in rdms]
    todo = [k for k, r in enumerate(out) if r is None]
    for k, item in zip(todo, PresortedRDM.batch([rdms[k] for k in todo])):
        out[k] = item
    return out


def bootstrap_rdm_correlation_grouped(model_rdms: Sequence, neural_rdms: Sequence, method: str,
                                      n_bootstrap: int = 1000, seed: int = 42,
                                      index_sets: Optional[np.ndarray] = None,
                                      rng: Optional[np.random.RandomState] = None,
                                      replicate_range: Optional[Tuple[int, int]] = None) -> Dict[str, np.ndarray]:
    """Point estimates and bootstrap scores for every (model RDM, neural RDM) pair in one pass.

    Every pair sees the SAME index sets, which is what the reference does by re-seeding
    RandomState(42) per (region, subject) (evals.py:356).  Returns host arrays
    {"point": (nA, nB), "scores": (nA, nB, B)} with replicates in draw order.
    replicate_range=(lo, hi) computes only that slice of the draws (multi-GPU sharding).
    """
    method = method.lower()
    dev = _lib.require_cuda()
    both = _prepare_many(list(model_rdms) + list(neural_rdms), method)
    A, B = both[:len(model_rdms)], both[len(model_rdms):]
    n = A[0].n
    if index_sets is None:
        index_sets = draw_subsample_masks(n, n_bootstrap, seed, rng, replicate_range)
    elif replicate_range is not None:
        index_sets = index_sets[replicate_range[0]:replicate_range[1]]
    masks = build_masks(index_sets, n, dev, with_full=True)
    kept = None if isinstance(index_sets, HostMasks) or index_sets is None or len(index_sets) == 0 else int(np.asarray(index_sets).shape[1])
    out = compare_grouped(A, B, method, masks, kept_stimuli=kept if kept is not None else int(n * SUBSAMPLE_FRACTION)).cpu().numpy()
    return {"point": out[:, :, 0].copy(), "scores": out[:, :, 1:].copy()}


def bootstrap_rdm_correlation(model_rdm, neural_rdm, method: str, n_bootstrap: int = 1000, seed: int = 42):
    """(score, ci_low, ci_high, scores) -- the reference's helper of the same name
    (run_reconstruction.py:137-158) and its inlined twins (evals.py:355-373, :506-522)."""
    res = bootstrap_rdm_correlation_grouped([model_rdm], [neural_rdm], method, n_bootstrap, seed)
    scores = res["scores"][0, 0]
    ci_low, ci_high = percentile_ci(scores)
    return (float(res["point"][0, 0]), float(ci_low), float(ci_high),
            scores.tolist())


def _flat(a: torch.Tensor) -> torch.Tensor:
    return a.flatten(start_dim=1) if a.ndim > 2 else a


def compute_rsa(cfg: Dict, selection: "AlignmentData", evaluation: "AlignmentData", n_select: Optional[int] = None,
                bootstrap: bool = True, n_bootstrap: int = 1000, seed: int = 42, verbose: bool = False,
                re_extract_fn=None) -> List[Dict]:
    """Train/test RSA (reference: rsa.py:132-281): pick the best layer on `selection`,
    score it on `evaluation`, optional 90 %-subsample bootstrap CI.

    One RandomState(seed) feeds the optional selection subsample first and the bootstrap draws
    afterwards, exactly as the reference consumes it (rsa.py:169,176,249).
    """
    method = cfg.get("compare_method", "spearman").lower()
    rng = np.random.RandomState(seed)
    n_train = selection.neural.size(0)
    n_test = evaluation.neural.size(0)
    if n_select is not None and n_select < n_train:
        sel_idx = rng.choice(n_train, size=n_select, replace=False)
    else:
        sel_idx = np.arange(n_train)
    if verbose:
        print(f"Train/test RSA: {n_train} train, {n_test} test, {len(sel_idx)} for layer selection; "
              f"Pearson RDMs, {method.capitalize()} comparison")

    # 1. layer selection: all layers against the neural RDM in one grouped device call
    if method not in _METHODS:
        raise ValueError("correlation must be 'Pearson', 'Spearman', or 'Kendall'")
    dev = _lib.require_cuda()
    neural_sel = _prepare(compute_rdm(selection.neural[sel_idx].to(dev)), method)
    names = list(selection.activations.keys())
    layer_rdms = _prepare_many([compute_rdm(_flat(selection.activations[k][sel_idx]).to(dev)) for k in names], method)
    full = build_masks(None, len(sel_idx), dev, with_full=True)
    sel = compare_grouped(layer_rdms, [neural_sel], method, full)[:, 0, 0].cpu().numpy()
    selection_scores, best_layer, best_score = [], None, -float("inf")
    for name, s in zip(names, sel):
        s = float(s)
        selection_scores.append({"layer": name, "score": s})
        if verbose:
            print(f"  [select] {name:<15} RSA = {s:.4f}")
        if s > best_score:
            best_score, best_layer = s, name
    del layer_rdms

    # 2. evaluation RDMs (exact re-extraction if a callback is supplied, rsa.py:212-218)
    if re_extract_fn is not None:
        exact_acts, _ = re_extract_fn(best_layer, evaluation.stimulus_ids)
        test_acts = _flat(exact_acts)
    else:
        test_acts = _flat(evaluation.activations[best_layer])
    neural_rdm = compute_rdm(evaluation.neural.to(dev))
    model_rdm = compute_rdm(test_acts.to(dev))

    # 3. point estimate + bootstrap in one grouped call (mask 0 = all stimuli)
    idx = None
    hook = _observed_compare()
    if hook is not None:
        # Someone wrapped compute_rdm_correlation (the reference's own call-structure test does,
        # tests/test_rsa_bootstrap.py:519-546): keep the reference's observable call sequence -- one call for the
        # point estimate and one per replicate on the sub-indexed RDMs (rsa.py:223, :248-256).
        point = hook(model_rdm, neural_rdm, correlation=method.capitalize())
        scores = np.empty(n_bootstrap if bootstrap else 0, dtype=np.float64)
        if bootstrap:
            draws = draw_index_sets(n_test, n_bootstrap, rng=rng)
            for b in range(n_bootstrap):
                sub = torch.from_numpy(draws[b].astype(np.int64)).to(model_rdm.device)
                scores[b] = hook(model_rdm[sub][:, sub], neural_rdm[sub][:, sub], correlation=method.capitalize())
    elif bootstrap:
        idx = draw_subsample_masks(n_test, n_bootstrap, rng=rng)
    if hook is not None:
        pass
    elif n_test <= 1:
        point, scores = float("nan"), np.full(n_bootstrap if bootstrap else 0, np.nan)
    else:
        masks = build_masks(idx, n_test, dev, with_full=True)
        out = compare_grouped([_prepare(model_rdm, method)], [_prepare(neural_rdm, method)], method, masks)[0, 0].cpu().numpy()
        point, scores = float(out[0]), out[1:]
    ci_low = ci_high = None
    result = {"layer": best_layer, "compare_method": method, "score": point, "ci_low": None, "ci_high": None,
              "analysis": "rsa", "layer_selection_scores": selection_scores}
    if bootstrap:
        ci_low, ci_high = (float(v) for v in percentile_ci(scores))
        result["ci_low"], result["ci_high"] = ci_low, ci_high
        result["bootstrap_scores"] = scores.tolist()
    if verbose:
        msg = f"  {method.capitalize():<10}| {best_layer} = {point:.4f}"
        if bootstrap:
            msg += f"  [95% CI: {ci_low:.4f}, {ci_high:.4f}]"
        print(msg)
    return [result]


def _observed_compare():
    """The callable to route per-replicate comparisons through when `compute_rdm_correlation` has been wrapped or
    replaced by the caller -- on this module or on the reference module install() patched -- else None (grouped path)."""
    import sys
    mine = globals().get("compute_rdm_correlation")
    if mine is not _COMPARE_IMPL:
        return mine
    ref = sys.modules.get("visreps.analysis.rsa")
    theirs = getattr(ref, "compute_rdm_correlation", None) if ref is not None else None
    if theirs is not None and theirs is not _COMPARE_IMPL and ("visreps.analysis.rsa", "compute_rdm_correlation") in _installed():
        return theirs
    return None


def _installed():
    import sys
    mod = sys.modules.get("visreps_b200.install")
    return mod._saved if mod is not None else {}


def _concept_average_exact(raw_acts: torch.Tensor, raw_ids, data) -> torch.Tensor:
    """Per-concept mean of per-image activations, in data.stimulus_ids order (reference: rsa.py:284-305).

    Producer-side glue: stays in PyTorch, one index_add instead of a Python loop."""
    where = {str(k): i for i, k in enumerate(raw_ids)}
    rows, owner = [], []
    for c, concept in enumerate(data.stimulus_ids):
        for sid in data.concept_image_ids[concept]:
            if sid in where:
                rows.append(where[sid])
                owner.append(c)
    n_c = len(data.stimulus_ids)
    out = torch.zeros((n_c, raw_acts.size(1)), dtype=torch.float32, device=raw_acts.device)
    if rows:
        r = torch.as_tensor(rows, device=raw_acts.device)
        o = torch.as_tensor(owner, device=raw_acts.device)
        # gather + accumulate in image blocks of <= 256 MB so the fp32 copy of the gathered rows stays bounded
        # (THINGS: 26k images x up to 290,400 features would otherwise be a second 30 GB copy)
        step = max(1, (256 << 20) // max(4, 4 * int(raw_acts.size(1))))
        for k0 in range(0, len(rows), step):
            out.index_add_(0, o[k0:k0 + step], raw_acts[r[k0:k0 + step]].float())
        cnt = torch.bincount(o, minlength=n_c).clamp_(min=1).to(out.dtype)
        out /= cnt[:, None]
    return out.to(raw_acts.dtype)


def subsample_rank(rdm, masks: torch.Tensor):
    """Centred tie-averaged ranks of every masked subsample of one RDM's upper triangle.

    Returns (dev, ss, nkept): dev is (R, N) int32 with dev = 2*rank - (N_r + 1) for kept pairs and
    0 elsewhere (canonical triu order), ss = sum(dev**2) per replicate, nkept = N_r.
    With the all-ones mask, dev + N + 1 is 2 * scipy.stats.rankdata(triu_vector, 'average').
    """
    lib = _lib.load()
    p = rdm if isinstance(rdm, PresortedRDM) else PresortedRDM(rdm)
    dev_ = masks.device
    R = int(masks.shape[0])
    stride = lib.vr_dev_stride(p.n)
    out = torch.empty((R, stride), dtype=torch.int32, device=dev_)
    ss = torch.empty(R, dtype=torch.float64, device=dev_)
    nk = torch.empty(R, dtype=torch.int64, device=dev_)
    ws = _lib.workspace(lib.vr_subsample_rank_workspace_bytes(p.n, R), dev_)
    _lib.check(lib.vr_subsample_rank(p.blob.data_ptr(), p.n, masks.data_ptr(), R, out.data_ptr(), ss.data_ptr(),
                                     nk.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "vr_subsample_rank")
    N = p.n * (p.n - 1) // 2
    return out[:, :N], ss, nk


def _rdm_device(x: torch.Tensor, dev: torch.device, correction: float, flags: torch.Tensor, slot: int,
                ws_slot: str = "main") -> torch.Tensor:
    """compute_rdm without the host round trip for the zero-variance count (collected in `flags`)."""
    lib = _lib.load()
    x = _as_cuda_f32_rows(_flat(x) if x.ndim > 2 else x, dev)
    n, d = int(x.size(0)), int(x.size(1))
    rdm = torch.empty((n, n), dtype=torch.float32, device=dev)
    ws = _lib.workspace(lib.vr_rdm_workspace_bytes(n, d), dev, ws_slot)
    _lib.check(lib.vr_rdm_pearson_f32(x.data_ptr(), n, d, x.stride(0), float(correction), rdm.data_ptr(),
                                      flags[slot:slot + 1].data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()),
               "vr_rdm_pearson_f32")
    return rdm


RDM_OVERLAP = True        # False: one stream, one matrix after the other (bench.py's per-kernel breakdown is taken that way)
_rdm_streams: Dict[int, Tuple[torch.cuda.Stream, torch.cuda.Stream]] = {}


def _rdm_device_many(xs: Sequence[torch.Tensor], dev: torch.device, correction: float, flags: torch.Tensor,
                     slots: Sequence[int]) -> List[torch.Tensor]:
    """RDMs of device-resident inputs, alternating between two streams with a scratch buffer each: the row-centring pass of
    one matrix (HBM bound, no shared memory) shares the SMs with the tensor-core SYRK of the previous one (one CTA per SM,
    bound by the tensor pipe) instead of waiting for it."""
    if len(xs) < 2 or not RDM_OVERLAP:
        return [_rdm_device(x, dev, correction, flags, k) for x, k in zip(xs, slots)]
    cur = torch.cuda.current_stream()
    pair = _rdm_streams.get(dev.index)
    if pair is None:
        pair = _rdm_streams[dev.index] = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
    for st in pair:
        st.wait_stream(cur)
    out = []
    for pos, (x, k) in enumerate(zip(xs, slots)):
        with torch.cuda.stream(pair[pos & 1]):
            rdm = _rdm_device(x, dev, correction, flags, k, ws_slot=f"rdm{pos & 1}")
        rdm.record_stream(cur)
        if isinstance(x, torch.Tensor) and x.is_cuda:
            x.record_stream(pair[pos & 1])
        out.append(rdm)
    for st in pair:
        cur.wait_stream(st)
    return out


class RankContext:
    """Everything the ranking kernels derive from the subsample masks alone (library-opaque), built once
    per job and shared by every RDM that is ranked against those masks."""

    def __init__(self, masks: torch.Tensor, n: int):
        lib = _lib.load()
        self.n, self.R, self.masks = n, int(masks.shape[0]), masks
        self.buf = torch.empty(lib.vr_rank_context_bytes(n, self.R), dtype=torch.uint8, device=masks.device)
        _lib.check(lib.vr_rank_context_build(masks.data_ptr(), n, self.R, self.buf.data_ptr(), self.buf.numel(),
                                             _lib.stream_ptr()), "vr_rank_context_build")

    def materialise(self, items: Sequence[PresortedRDM]):
        """dev (nV, R, stride) int32 and ss (nV, R) fp64 of the given presorted RDMs."""
        lib = _lib.load()
        dev = self.buf.device
        nV = len(items)
        stride = lib.vr_dev_stride(self.n)
        out = torch.empty((nV, self.R, stride), dtype=torch.int32, device=dev)
        ss = torch.empty((nV, self.R), dtype=torch.float64, device=dev)
        ws = _lib.workspace(lib.vr_rank_workspace_bytes(self.n, nV, self.R), dev, "rank")
        _lib.check(lib.vr_rank_materialise(_ptr_array([p.blob for p in items]), nV, self.n, self.buf.data_ptr(), self.R,
                                           out.data_ptr(), ss.data_ptr(), None, ws.data_ptr(), ws.numel(),
                                           _lib.stream_ptr()), "vr_rank_materialise")
        return out, ss

    def fused(self, items: Sequence[PresortedRDM], mat_dev: torch.Tensor, mat_ss: torch.Tensor, out: torch.Tensor,
              out_offset: int, stride_fused: int, stride_mat: int) -> None:
        """Rank `items` and write their Spearman coefficients against the materialised vectors into `out`
        (flat fp64) at out_offset + f * stride_fused + m * stride_mat + r."""
        lib = _lib.load()
        dev = self.buf.device
        ws = _lib.workspace(lib.vr_rank_workspace_bytes(self.n, len(items), self.R), dev, "rank")
        _lib.check(lib.vr_rank_fused(_ptr_array([p.blob for p in items]), len(items), self.n, self.buf.data_ptr(), self.R,
                                     mat_dev.data_ptr(), mat_ss.data_ptr(), int(mat_dev.shape[0]),
                                     out.data_ptr() + 8 * out_offset, stride_fused, stride_mat, ws.data_ptr(), ws.numel(),
                                     _lib.stream_ptr()), "vr_rank_fused")


def _spearman_pipeline(inputs_a: Sequence[torch.Tensor], inputs_b: Sequence[torch.Tensor], index_sets: Optional[np.ndarray],
                       draw, n_draws: int, correction: float, dev: torch.device):
    """Pearson RDMs + grouped Spearman for host or device inputs, with host->device copies, RDM builds and
    ranking overlapped.  Returns (out (nA, nB, R) fp64 device tensor, zero-variance flags tensor).

    The side with fewer matrices is "materialised" first (its centred rank vectors are kept in HBM); every
    matrix of the other side is then copied (smallest first), turned into an RDM, presorted and ranked
    against the materialised vectors as soon as its copy lands, so only the last matrix's work is exposed
    after the final copy.
    """
    lib = _lib.load()
    nA, nB = len(inputs_a), len(inputs_b)
    n = int(inputs_a[0].shape[0])
    mat_is_a = nA < nB                                   # ties: the neural side is materialised
    mats = list(inputs_a if mat_is_a else inputs_b)
    fus = list(inputs_b if mat_is_a else inputs_a)
    # Copy order: smallest first (the GPU gets work at once), but the third-largest matrix LAST: what is exposed after
    # the final copy is that matrix's RDM build plus its ranking, and the ranking costs the same for every matrix, so
    # the last one should be cheap to build -- yet large enough that the two biggest builds are over by then.
    order = sorted(range(len(fus)), key=lambda k: fus[k].numel())     # smallest copy first
    if len(order) >= 4 and any(f.device != dev for f in fus):
        order.append(order.pop(-3))
    # size the scratch for the largest matrix now, so no allocation (and no implicit sync) happens mid-pipeline
    _lib.workspace(max(lib.vr_rdm_workspace_bytes(n, int(np.prod(x.shape[1:]))) for x in mats + fus), dev)
    compute = torch.cuda.current_stream()
    copier = _copy_stream(dev)
    copier.wait_stream(compute)

    def stage(x):                                        # enqueue the H2D copy (if any) on the copy stream
        x = _flat(x) if x.ndim > 2 else x
        if x.device == dev:
            return x, None
        with torch.cuda.stream(copier):
            xd = x.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copier)
            _mark(f"copied {tuple(x.shape)}", copier)
        xd.record_stream(compute)
        return xd, ev

    _mark("start")

    staged_m = [stage(x) for x in mats]
    staged_f = {k: stage(fus[k]) for k in order}
    flags = torch.zeros(nA + nB, dtype=torch.int32, device=dev)
    gmax = lib.vr_rank_fuse_max()

    def rdm_of(staged, slot):
        xd, ev = staged
        if ev is not None:
            compute.wait_event(ev)
        return _rdm_device(xd, dev, correction, flags, slot)

    slot_m = lambda k: k if mat_is_a else nA + k
    slot_f = lambda k: nA + k if mat_is_a else k
    fus_on_device = all(ev is None for _, ev in staged_f.values())
    # Work that does not need the masks is enqueued before the host draws the index sets: the
    # materialised side, and as many of the fused side's matrices (smallest first) as will have landed
    # by the time the draw is done (host draw ~8.5 ns per index, copies ~50 GB/s).  The compute stream
    # only waits on copy events, never the host.
    f_items: Dict[int, PresortedRDM] = {}
    if fus_on_device and len(mats) + len(fus) <= _MAX_GROUP * 2:
        # everything is here already: one batched presort for both sides (half the sort launches)
        for _xd, ev in list(staged_m) + [staged_f[k] for k in range(len(fus))]:
            if ev is not None:
                compute.wait_event(ev)
        both = PresortedRDM.batch(_rdm_device_many([st[0] for st in staged_m] + [staged_f[k][0] for k in range(len(fus))], dev, correction, flags,
                                                   [slot_m(k) for k in range(len(staged_m))] + [slot_f(k) for k in range(len(fus))]))
        m_items = both[:len(mats)]
        f_items = dict(enumerate(both[len(mats):]))
    else:
        m_items = PresortedRDM.batch([rdm_of(st, slot_m(k)) for k, st in enumerate(staged_m)])
    if f_items:
        pass
    elif fus_on_device:
        for k, it in enumerate(PresortedRDM.batch([rdm_of(staged_f[k], slot_f(k)) for k in range(len(fus))])):
            f_items[k] = it
    elif index_sets is None:
        budget = 8.5e-9 * n * max(1, n_draws) * 50e9
        sent = sum(x.numel() * 4 for x in mats)
        for k in order:
            sent += fus[k].numel() * 4
            if sent > budget:
                break
            f_items[k] = PresortedRDM(rdm_of(staged_f[k], slot_f(k)))
    # ... while the host replays the reference's RNG stream and the copies fly
    if index_sets is None:
        index_sets = draw()
    _mark("drawn")
    masks = build_masks(index_sets, n, dev, with_full=True)
    R = int(masks.shape[0])
    out = torch.empty((nA, nB, R), dtype=torch.float64, device=dev)
    if SPEARMAN_ENGINE == "sweep" and (fus_on_device or n > _LEGACY_SMEM_TABLE_MAX_N):
        # One job holds the count tables of every RDM; the materialised side's tables are built first, every other
        # matrix is added and swept against them as soon as its copy has landed (its lookups are shared by all pairs
        # of the sweep), so only the last matrix's RDM + tables + sweep are exposed after the final copy.
        job = SpearmanJob(masks, n, nA + nB)
        n_m = len(m_items)
        m_slots = list(range(n_m))
        job.add(m_items, 0)
        _mark("tables of the materialised side")
        flat = out.view(-1)

        def sweep_fused(ks):                              # fused matrices ks (consecutive slots) x all materialised ones
            f_slots = [n_m + k for k in ks]
            if mat_is_a:
                job.sweep(m_slots, f_slots, flat, ks[0] * R, nB * R, R)
            else:
                job.sweep(f_slots, m_slots, flat, ks[0] * nB * R, nB * R, R)

        def add_and_sweep(ks):
            ks = sorted(ks)
            runs, cur = [], [ks[0]]
            for k in ks[1:]:
                if k == cur[-1] + 1 and len(cur) < 8:
                    cur.append(k)
                else:
                    runs.append(cur)
                    cur = [k]
            runs.append(cur)
            for run in runs:                               # consecutive slots: one table build, one sweep
                job.add([f_items[k] for k in run], n_m + run[0])
                sweep_fused(run)

        if fus_on_device:
            add_and_sweep(list(range(len(fus))))
        else:
            ready = [k for k in order if k in f_items]
            if ready:
                add_and_sweep(ready)
                _mark(f"swept batch {sorted(ready)}")
            for k in order:
                if k in f_items:
                    continue
                f_items[k] = PresortedRDM(rdm_of(staged_f[k], slot_f(k)))
                add_and_sweep([k])
                _mark(f"swept {k}")
        _mark("done")
        return out, flags
    ctx = RankContext(masks, n)
    _mark("context")

    # materialised side
    groups = []
    for g0 in range(0, len(m_items), gmax):
        mdev, mss = ctx.materialise(m_items[g0:g0 + gmax])
        groups.append((g0, mdev, mss))
    _mark("materialised")
    # fused side
    stride_m = nB * R if mat_is_a else R
    stride_f = R if mat_is_a else nB * R
    if fus_on_device:                                    # nothing to overlap with: one batched pass
        batch = [f_items[k] for k in range(len(fus))]
        for g0, mdev, mss in groups:
            ctx.fused(batch, mdev, mss, out, g0 * stride_m, stride_f, stride_m)
    else:
        ready = [k for k in order if k in f_items]       # presorted before the draw: one batched call per group
        if ready:
            # fused vector k lives at k * stride_f: the batch is addressed through its first member and the
            # (uniform) distance between consecutive members only when they are contiguous; otherwise one by one
            contiguous = all(b - a == 1 for a, b in zip(sorted(ready), sorted(ready)[1:]))
            if contiguous and len(ready) > 1:
                ks = sorted(ready)
                for g0, mdev, mss in groups:
                    ctx.fused([f_items[k] for k in ks], mdev, mss, out, ks[0] * stride_f + g0 * stride_m, stride_f, stride_m)
            else:
                for k in ready:
                    for g0, mdev, mss in groups:
                        ctx.fused([f_items[k]], mdev, mss, out, k * stride_f + g0 * stride_m, stride_f, stride_m)
            _mark(f"fused batch {sorted(ready)}")
        for k in order:
            if k in f_items:
                continue
            item = PresortedRDM(rdm_of(staged_f[k], slot_f(k)))
            for g0, mdev, mss in groups:
                ctx.fused([item], mdev, mss, out, k * stride_f + g0 * stride_m, stride_f, stride_m)
            _mark(f"fused {k}")
    _mark("done")
    return out, flags


_copy_streams: Dict[int, torch.cuda.Stream] = {}
TRACE: Optional[list] = None      # set to [] to collect (label, stream-event, host-seconds) marks of the pipeline


def _mark(label: str, stream=None) -> None:
    if TRACE is not None:
        import time
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(stream if stream is not None else torch.cuda.current_stream())
        TRACE.append((label, ev, time.perf_counter()))


def _copy_stream(dev: torch.device) -> torch.cuda.Stream:
    s = _copy_streams.get(dev.index)
    if s is None:
        s = _copy_streams[dev.index] = torch.cuda.Stream(device=dev)
    return s


def rsa_grouped(model_features: Sequence[torch.Tensor], neural_responses: Sequence[torch.Tensor],
                method: str = "spearman", bootstrap: bool = True, n_bootstrap: int = 1000, seed: int = 42,
                correction: float = 1e-12, replicate_range: Optional[Tuple[int, int]] = None,
                index_sets: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """The scoring block of `_eval_rsa` (visreps/evals.py:333-398) for many layers and many
    (region, subject) response matrices at once, with one host<->device round trip.

    model_features: (n, d_l) activations per layer (host or device; >2-D inputs are flattened like
    evals.py:321); neural_responses: (n, V_k) per ROI/subject.  For every (layer, ROI) pair:
    Pearson RDMs -> point estimate (evals.py:347) -> `n_bootstrap` 90 %-subsample replicates with the
    index sets of RandomState(seed), identical for every pair as in the reference (evals.py:356) ->
    percentile CI (evals.py:371-372).
    Returns host arrays: point (nA, nB), scores (nA, nB, B) in draw order, ci_low / ci_high (nA, nB).
    """
    with _no_gc():
        return _rsa_grouped(model_features, neural_responses, method, bootstrap, n_bootstrap, seed, correction,
                            replicate_range, index_sets)


def _rsa_grouped(model_features, neural_responses, method, bootstrap, n_bootstrap, seed, correction, replicate_range, index_sets):
    method = method.lower()
    if method not in _METHODS:
        raise ValueError("correlation must be 'Pearson', 'Spearman', or 'Kendall'")
    dev = _lib.require_cuda()
    _mark("enter")
    n = int(model_features[0].shape[0])
    if any(int(t.shape[0]) != n for t in list(model_features) + list(neural_responses)):
        raise ValueError("all inputs must have the same number of stimuli")
    nA, nB = len(model_features), len(neural_responses)

    def draw():
        if not bootstrap:
            return None
        return draw_subsample_masks(n, n_bootstrap, seed, replicate_range=replicate_range)

    if index_sets is not None and replicate_range is not None:
        index_sets = index_sets[replicate_range[0]:replicate_range[1]]
    if not bootstrap:
        index_sets = None
    if n <= 1:
        b = 0 if not bootstrap else (len(index_sets) if index_sets is not None else n_bootstrap)
        return {"point": np.full((nA, nB), np.nan), "scores": np.full((nA, nB, b), np.nan),
                "ci_low": np.full((nA, nB), np.nan), "ci_high": np.full((nA, nB), np.nan)}
    if method == "spearman" and _pipeline_fits(n, nA, nB, (n_bootstrap if bootstrap else 0) + 1, dev):
        out_dev, flags = _spearman_pipeline(model_features, neural_responses, index_sets, draw,
                                            n_bootstrap if bootstrap else 0, correction, dev)
    else:
        flags = torch.zeros(nA + nB, dtype=torch.int32, device=dev)
        every = list(model_features) + list(neural_responses)
        if all(isinstance(x, torch.Tensor) and x.is_cuda for x in every):
            rdms = _rdm_device_many(every, dev, correction, flags, list(range(len(every))))
        else:
            rdms = [_rdm_device(x, dev, correction, flags, k) for k, x in enumerate(every)]
        prepared = _prepare_many(rdms, method)
        if index_sets is None:
            index_sets = draw()
        masks = build_masks(index_sets, n, dev, with_full=True)
        out_dev = compare_grouped(prepared[:nA], prepared[nA:], method, masks)
    # the CI bounds of evals.py:371-372 are taken on the device too (a sort of 64 x 1000 scores on the host is 0.5 ms after
    # the last kernel, with nothing to hide behind)
    ci_dev = percentile_ci_device(out_dev[:, :, 1:]) if (bootstrap and out_dev.shape[2] > 1) else None
    host = torch.empty(out_dev.shape, dtype=torch.float64, pin_memory=True)
    host.copy_(out_dev, non_blocking=True)
    ci_host = None
    if ci_dev is not None:
        ci_host = torch.empty(ci_dev.shape, dtype=torch.float64, pin_memory=True)
        ci_host.copy_(ci_dev, non_blocking=True)
    flags_host = flags.cpu()            # synchronises the stream: results are complete after this
    torch.cuda.current_stream().synchronize()
    _mark("synced")
    nz = int(flags_host.sum())
    if nz:
        logger.warning("%d rows have ~zero variance across %d RDMs", nz, nA + nB)
    out = host.numpy()
    res = {"point": out[:, :, 0].copy(), "scores": out[:, :, 1:].copy()}
    if ci_host is not None:
        res["ci_low"], res["ci_high"] = ci_host[0].numpy().copy(), ci_host[1].numpy().copy()
    elif bootstrap and res["scores"].shape[2] > 0:
        res["ci_low"], res["ci_high"] = percentile_ci(res["scores"])
    _mark("exit")
    return res


def _pipeline_fits(n: int, nA: int, nB: int, R: int, dev: torch.device) -> bool:
    """The pipelined path keeps all R replicates of the materialised side resident; fall back to the
    replicate-chunked one-call path when that would not fit comfortably."""
    lib = _lib.load()
    if SPEARMAN_ENGINE == "sweep":
        key = ("sweep", n, nA, nB, R, dev.index)
        hit = _fit_cache.get(key)
        if hit is None:
            free, _total = torch.cuda.mem_get_info(dev)
            cached = torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
            hit = _fit_cache[key] = SpearmanJob.bytes_needed(n, nA + nB, R) < 0.5 * (free + cached)
        return hit
    n_mat = min(nA, nB)
    # vr_rank_fused addresses a materialised group (<= vr_rank_fuse_max() vectors x R replicates) with 32-bit int4
    # offsets: 2^34 elements.  Beyond that the one-call path chunks the replicates itself.
    if min(n_mat, lib.vr_rank_fuse_max()) * R * lib.vr_dev_stride(n) >= (1 << 34):
        return False
    need = n_mat * R * (lib.vr_dev_stride(n) * 4 + 8) + lib.vr_rank_context_bytes(n, R) + \
        lib.vr_rank_workspace_bytes(n, max(min(n_mat, lib.vr_rank_fuse_max()), max(nA, nB)), R)
    key = (n, nA, nB, R, dev.index)
    hit = _fit_cache.get(key)
    if hit is None:      # cudaMemGetInfo costs 0.2-1.5 ms: asked once per problem shape, not once per call
        free, _total = torch.cuda.mem_get_info(dev)
        cached = torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
        hit = _fit_cache[key] = need < 0.6 * (free + cached)
    return hit


_fit_cache: Dict[tuple, object] = {}


class _no_gc:
    """The enqueue phase of a step is a few hundred microseconds of host work that gates ~25 ms of device
    work; a generational collection landing in it (tens of ms on a large heap) stalls the whole step.
    Collection is deferred until the step's results are back."""

    def __enter__(self):
        import gc
        self._was = gc.isenabled()
        gc.disable()

    def __exit__(self, *exc):
        if self._was:
            import gc
            gc.enable()
        return False
