"""Rebind the reference's RSA entry points to this package (drop-in for the hot path only).

The reference binds by name: `from visreps.analysis.rsa import compute_rdm, compute_rdm_correlation`
(visreps/evals.py:25) and `from visreps.analysis.rsa import compute_rsa` (visreps/analysis/alignment.py:8), and
`eval()` calls its helpers `_eval_rsa` (evals.py:222) and `_eval_rsa_nsd_synthetic` (:442) through the module
globals.  install() patches whichever of those modules are importable / already imported:

  * the numerics: rsa.{compute_rdm, compute_rdm_correlation, compute_rsa, _rank, _concept_average_exact},
    evals.{compute_rdm, compute_rdm_correlation}, alignment.compute_rsa;
  * the loops: evals.{_eval_rsa, _eval_rsa_nsd_synthetic} -> visreps_b200.evals (same signatures, same DataFrame;
    layer selection grouped across pairs, all pairs' replicates in one device job);
  * optionally the producer hand-off (device_producer=True): models.utils.{get_activations, extract_single_layer}
    -> visreps_b200.producer (activations stay on the GPU, models/utils.py:344,391);
  * optionally the results sink (batched_sink=True): one SQLite transaction per job (visreps/utils.py:381-458).

A target that cannot be imported is reported with a warning (and raised with strict=True), never skipped silently.
"""
from __future__ import annotations

import importlib
import sys
import warnings

_TARGETS = {
    "visreps.analysis.rsa": ("compute_rdm", "compute_rdm_correlation", "compute_rsa", "_rank", "_concept_average_exact"),
    "visreps.evals": ("compute_rdm", "compute_rdm_correlation", "_concept_average_exact", "_eval_rsa", "_eval_rsa_nsd_synthetic"),
    "visreps.analysis.alignment": ("compute_rsa",),
    # secondary callers: scripts, patched only if they are already imported
    "experiments.reconstruction_analysis.run_reconstruction": ("compute_rdm", "compute_rdm_correlation", "bootstrap_rdm_correlation"),
    "experiments.binary_pc_rsa.main": ("compute_rdm", "_compare_rdms"),
}
_PRODUCER_TARGETS = {"visreps.models.utils": ("get_activations", "extract_single_layer")}
_SCRIPTS = ("experiments.reconstruction_analysis.run_reconstruction", "experiments.binary_pc_rsa.main")
_EVALS_SYMBOLS = ("_compare_rdms", "_eval_rsa", "_eval_rsa_nsd_synthetic")
_saved = {}


class InstallWarning(UserWarning):
    """A reference module that should have been patched could not be imported."""


def _ours(sym: str, mod_name: str):
    from . import rsa, evals, producer
    if mod_name in _PRODUCER_TARGETS:
        return getattr(producer, sym)
    return getattr(evals, sym) if sym in _EVALS_SYMBOLS else getattr(rsa, sym)


def install(strict: bool = False, device_producer: bool = False, batched_sink: bool = False):
    """Patch the reference modules in place; returns the list of (module, symbol) rebound."""
    done = []
    targets = dict(_TARGETS)
    if device_producer:
        targets.update(_PRODUCER_TARGETS)
    for mod_name, symbols in targets.items():
        mod = sys.modules.get(mod_name)
        if mod is None:
            if mod_name in _SCRIPTS:      # never import (i.e. run) an experiment script just to patch it
                continue
            try:
                mod = importlib.import_module(mod_name)
            except Exception as exc:
                if strict:
                    raise
                warnings.warn(f"visreps_b200.install(): could not import {mod_name} ({type(exc).__name__}: {exc}); "
                              f"{', '.join(symbols)} stay on the reference CPU path there", InstallWarning, stacklevel=2)
                continue
        for sym in symbols:
            if hasattr(mod, sym):
                _saved.setdefault((mod_name, sym), getattr(mod, sym))
                setattr(mod, sym, _ours(sym, mod_name))
                done.append((mod_name, sym))
    if batched_sink:
        from . import evals, results_sink
        evals.RESULTS_SINK = results_sink.save_results_batch
    return done


def uninstall():
    for (mod_name, sym), fn in list(_saved.items()):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            setattr(mod, sym, fn)
        del _saved[(mod_name, sym)]
    from . import evals
    evals.RESULTS_SINK = None
