"""visreps_b200: B200-native (sm_100a) implementation of the visreps RSA hot path.

Drop-in for visreps/analysis/rsa.py's entry points; see DESIGN.md and INTEGRATION.md.
"""
from .rsa import (  # noqa: F401
    PresortedRDM,
    bootstrap_rdm_correlation,
    bootstrap_rdm_correlation_grouped,
    build_masks,
    compare_grouped,
    compute_rdm,
    compute_rdm_correlation,
    compute_rsa,
    draw_index_sets,
    rsa_grouped,
    _rank,
    _concept_average_exact,
)
from .alignment import AlignmentData, compute_traintest_alignment, _align_stimulus_level  # noqa: F401
from .install import install, uninstall  # noqa: F401
from . import evals  # noqa: F401
from . import encoding  # noqa: F401

__all__ = ["compute_rdm", "compute_rdm_correlation", "compute_rsa", "bootstrap_rdm_correlation",
           "bootstrap_rdm_correlation_grouped", "rsa_grouped", "AlignmentData", "compute_traintest_alignment", "install", "uninstall"]
