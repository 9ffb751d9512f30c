"""Rebind the reference's RSA entry points to this package (drop-in for the hot path only).

The reference binds by name: `from visreps.analysis.rsa import compute_rdm, compute_rdm_correlation`
(visreps/evals.py:25) and `from visreps.analysis.rsa import compute_rsa` (visreps/analysis/alignment.py:8).
install() patches whichever of those modules are importable / already imported.  The bootstrap loops
inlined in visreps/evals.py:355-373 and :506-522 call compute_rdm_correlation once per replicate; the
one-line change that turns them into a single device call is shown in INTEGRATION.md.
"""
from __future__ import annotations

import importlib
import sys

_TARGETS = {
    "visreps.analysis.rsa": ("compute_rdm", "compute_rdm_correlation", "compute_rsa", "_rank", "_concept_average_exact"),
    "visreps.evals": ("compute_rdm", "compute_rdm_correlation"),
    "visreps.analysis.alignment": ("compute_rsa",),
    # secondary callers: scripts, patched only if they are already imported
    "experiments.reconstruction_analysis.run_reconstruction": ("compute_rdm", "compute_rdm_correlation", "bootstrap_rdm_correlation"),
    "experiments.binary_pc_rsa.main": ("compute_rdm", "_compare_rdms"),
}
_SCRIPTS = ("experiments.reconstruction_analysis.run_reconstruction", "experiments.binary_pc_rsa.main")
_saved = {}


def _ours(sym: str):
    from . import rsa, evals
    return getattr(evals, sym) if sym == "_compare_rdms" else getattr(rsa, sym)


def install(strict: bool = False):
    """Patch the reference modules in place; returns the list of (module, symbol) rebound."""
    done = []
    for mod_name, symbols in _TARGETS.items():
        mod = sys.modules.get(mod_name)
        if mod is None:
            if mod_name in _SCRIPTS:      # never import (i.e. run) an experiment script just to patch it
                continue
            try:
                mod = importlib.import_module(mod_name)
            except Exception:
                if strict:
                    raise
                continue
        for sym in symbols:
            if hasattr(mod, sym):
                _saved.setdefault((mod_name, sym), getattr(mod, sym))
                setattr(mod, sym, _ours(sym))
                done.append((mod_name, sym))
    return done


def uninstall():
    for (mod_name, sym), fn in list(_saved.items()):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            setattr(mod, sym, fn)
        del _saved[(mod_name, sym)]
