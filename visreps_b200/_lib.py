"""ctypes binding of the C-ABI library (include/visreps_b200.h).

The product path has no CPU implementation: if the shared library is missing, or CUDA is
not available, every entry point raises.  PyTorch is used only for device memory, streams
and torch.distributed.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvisreps_b200.so")

VR_OK, VR_ERR_INVALID, VR_ERR_WORKSPACE, VR_ERR_CUDA, VR_ERR_UNSUPPORTED = range(5)
ABI_VERSION = 3

# name -> (restype, argtypes); mirrors include/visreps_b200.h one to one
_vp, _i, _i64, _sz, _f = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float
SIGNATURES = {
    "vr_abi_version": (_i, []),
    "vr_last_error_string": (C.c_char_p, []),
    "vr_launch_count": (C.c_ulonglong, []),
    "vr_set_option": (_i, [C.c_char_p, _i]),
    "vr_get_option": (_i, [C.c_char_p, C.POINTER(C.c_int)]),
    "vr_profile_family_count": (_i, []),
    "vr_profile_family_name": (C.c_char_p, [_i]),
    "vr_profile_start": (_i, []),
    "vr_profile_stop": (_i, [_vp, _vp, _vp]),
    "vr_rdm_workspace_bytes": (_sz, [_i, _i64]),
    "vr_rdm_pearson_f32": (_i, [_vp, _i, _i64, _i64, _f, _vp, _vp, _vp, _sz, _vp]),
    "vr_rdm_pearson_triu_f32": (_i, [_vp, _i, _i64, _i64, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
    "vr_triu_pack_f32": (_i, [_vp, _i, _i64, _vp, _vp]),
    "vr_rdm_moments_f32": (_i, [_vp, _i, _i64, _i64, _vp, _vp, _vp]),
    "vr_rdm_shift_from_moments": (_i, [_vp, _i, _i64, _vp, _vp]),
    "vr_rdm_gram_f32": (_i, [_vp, _i, _i64, _i64, _vp, _vp, _i64, _vp, _sz, _vp]),
    "vr_rdm_finish_f32": (_i, [_vp, _i64, _i, _i64, _vp, _vp, _vp, _f, _vp, _vp, _vp]),
    "vr_rank_rows_workspace_bytes": (_sz, [_i, _i64]),
    "vr_rank_rows_ordinal_f32": (_i, [_vp, _i, _i64, _i64, _vp, _vp, _sz, _vp]),
    "vr_presort_bytes": (_sz, [_i]),
    "vr_presort_workspace_bytes": (_sz, [_i]),
    "vr_presort_triu_f32": (_i, [_vp, _i, _i64, _vp, _vp, _sz, _vp]),
    "vr_presort_batch_workspace_bytes": (_sz, [_i, _i]),
    "vr_presort_triu_batch_f32": (_i, [_vp, _i, _i, _i64, _vp, _vp, _sz, _vp]),
    "vr_host_draw_subsamples": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "vr_mask_words": (_i, [_i]),
    "vr_build_masks": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "vr_host_draw_subsample_masks": (_i, [_vp, C.POINTER(C.c_int), _i, _i, _i, _i, _i, _vp]),
    "vr_fetch_masks": (_i, [_vp, _i, _i, _vp, _vp]),
    "vr_dev_stride": (_i64, [_i]),
    "vr_subsample_rank_workspace_bytes": (_sz, [_i, _i]),
    "vr_subsample_rank": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "vr_rank_context_bytes": (_sz, [_i, _i]),
    "vr_rank_context_build": (_i, [_vp, _i, _i, _vp, _sz, _vp]),
    "vr_rank_workspace_bytes": (_sz, [_i, _i, _i]),
    "vr_rank_fuse_max": (_i, []),
    "vr_rank_materialise": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "vr_rank_fused": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _i64, _i64, _vp, _sz, _vp]),
    "vr_spearman_job_bytes": (_sz, [_i, _i, _i]),
    "vr_spearman_job_init": (_i, [_vp, _i, _i, _i, _vp, _sz, _vp]),
    "vr_spearman_job_add": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _vp]),
    "vr_spearman_sweep_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "vr_spearman_job_sweep": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _i64, _i64, _vp, _sz, _vp]),
    "vr_bootstrap_spearman_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "vr_bootstrap_spearman": (_i, [_vp, _i, _vp, _i, _i, _vp, _i, _vp, _vp, _sz, _vp]),
    "vr_bootstrap_column_corr_workspace_bytes": (_sz, [_i, _i, _i]),
    "vr_bootstrap_column_corr_f32": (_i, [_vp, _vp, _i, _i, _i64, _i64, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "vr_kendall_workspace_bytes": (_sz, [_i]),
    "vr_kendall_batch_workspace_bytes": (_sz, [_i, _i]),
    "vr_kendall_tau_a": (_i, [_vp, _vp, _i, _i64, _vp, _i, _vp, _vp, _sz, _vp]),
    "vr_bootstrap_pearson_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "vr_bootstrap_pearson": (_i, [_vp, _i, _vp, _i, _i, _i64, _vp, _i, _vp, _vp, _sz, _vp]),
    "vr_percentile_ci_max_replicates": (_i, []),
    "vr_percentile_ci_f64": (_i, [_vp, _i, _i, _i64, C.c_double, C.c_double, _vp, _vp, _vp]),
}


class VisrepsB200Error(RuntimeError):
    pass


_lock = threading.Lock()
_lib = None


def load(path: str = LIB_PATH) -> C.CDLL:
    """dlopen the library and type every exported symbol.  Fails loudly."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(path):
            raise VisrepsB200Error(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C visreps_b200/csrc`). There is no CPU fallback for this path.")
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the .so diverge
            fn.restype, fn.argtypes = res, args
        if lib.vr_abi_version() != ABI_VERSION:
            raise VisrepsB200Error(f"ABI version {lib.vr_abi_version()} != expected {ABI_VERSION}")
        _lib = lib
        return lib


def check(rc: int, what: str = "") -> None:
    if rc == VR_OK:
        return
    msg = load().vr_last_error_string().decode("utf-8", "replace")
    if rc == VR_ERR_INVALID:
        raise ValueError(f"{what}: {msg}")
    raise VisrepsB200Error(f"{what}: status {rc}: {msg}")


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise VisrepsB200Error("visreps_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


_workspaces = {}


def workspace(nbytes: int, device: torch.device, slot: str = "main") -> torch.Tensor:
    """Grow-only per-(device, slot) scratch buffer (uint8)."""
    key = (device.index, slot)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            del _workspaces[key]
            del buf
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def release_workspaces() -> None:
    _workspaces.clear()


def set_option(name: str, value: int) -> None:
    check(load().vr_set_option(name.encode(), int(value)), "vr_set_option")


def get_option(name: str) -> int:
    v = C.c_int(0)
    check(load().vr_get_option(name.encode(), C.byref(v)), "vr_get_option")
    return int(v.value)


def launch_count() -> int:
    return int(load().vr_launch_count())


def profile_start() -> None:
    check(load().vr_profile_start(), "vr_profile_start")


def profile_stop() -> dict:
    """{family: {"ms": device ms, "scopes": bracketed stages, "launches": kernel launches}}"""
    lib = load()
    k = lib.vr_profile_family_count()
    ms = (C.c_double * k)()
    sc = (C.c_ulonglong * k)()
    la = (C.c_ulonglong * k)()
    check(lib.vr_profile_stop(ms, sc, la), "vr_profile_stop")
    return {lib.vr_profile_family_name(i).decode(): {"ms": ms[i], "scopes": int(sc[i]), "launches": int(la[i])}
            for i in range(k)}
