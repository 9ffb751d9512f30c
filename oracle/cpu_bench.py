"""CPU timing of the reference path for bench.py (cpu_baseline leg and --impl reference).
TEST/BENCH INFRASTRUCTURE: the thing timed here is the oracle port of the reference code
(oracle/rsa_oracle.py: the same torch-CPU sgemm + scipy.stats.spearmanr calls the reference
makes), on the box's own host cores.  It is never the product path.

A bounded sample of the NSD-shape workload is timed and scaled linearly to the whole job:
  * all 16 RDMs of the workload are built once with torch CPU threads = cores          (measured)
  * replicates are independent and SciPy ranking is single-threaded, so the bootstrap loop
    (visreps/evals.py:361-369) is fanned over `cores` worker processes, each running
    `reps_per_worker` replicates of one (layer, ROI) pair                                  (measured)
  * job time = t_rdm + n_compares * (t_fan / (cores * reps_per_worker))                    (scaled)
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time
from typing import Dict

import numpy as np
import torch

from . import rsa_oracle as orc

_shared = {}


def usable_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def _worker(args):
    start, count = args
    torch.set_num_threads(1)
    a, b, idx = _shared["a"], _shared["b"], _shared["idx"]
    t0 = time.perf_counter()
    s = orc.bootstrap_compare_with_indices(a, b, "spearman", idx[start:start + count])
    return time.perf_counter() - t0, float(s.sum())


class ReferenceSampler:
    """Builds the workload's inputs on the CPU once, forks a worker pool, times bounded samples."""

    def __init__(self, n: int, layer_dims: Dict[str, int], roi_dims: Dict[str, int], n_bootstrap: int,
                 cores: int | None = None, reps_per_worker: int = 8, seed: int = 0):
        from visreps_b200 import synth  # data generator only (no CUDA, no kernels)
        self.n, self.n_bootstrap = n, n_bootstrap
        self.cores = cores or usable_cores()
        self.reps_per_worker = reps_per_worker
        self.layer_dims, self.roi_dims = dict(layer_dims), dict(roi_dims)
        torch.set_num_threads(self.cores)
        z = synth.latent(n, torch.device("cpu"), seed * 1000 + 1)
        # features are built lazily per RDM to bound host memory (conv1 alone is 774 MB)
        self._z, self._seed = z, seed
        small_layer = min(layer_dims, key=layer_dims.get)
        small_roi = min(roi_dims, key=roi_dims.get)
        xa = synth.layer_features(z, layer_dims[small_layer], seed * 1000 + 10)
        xb = synth.roi_responses(z, roi_dims[small_roi], 500)
        _shared["a"], _shared["b"] = orc.rdm(xa), orc.rdm(xb)
        _shared["idx"] = orc.draw_subsamples(n, self.cores * reps_per_worker, 42)
        self.pool = mp.get_context("fork").Pool(self.cores)
        self.pool.map(_worker, [(0, 1)] * self.cores)  # start-up + warm caches

    def close(self):
        self.pool.close()
        self.pool.join()

    def time_rdms(self) -> float:
        """Seconds to build every RDM of the workload with the reference arithmetic."""
        from visreps_b200 import synth
        total = 0.0
        for i, (k, d) in enumerate(self.layer_dims.items()):
            x = synth.layer_features(self._z, d, self._seed * 1000 + 10 + i)
            t0 = time.perf_counter()
            orc.rdm(x)
            total += time.perf_counter() - t0
            del x
        for i, (k, v) in enumerate(self.roi_dims.items()):
            x = synth.roi_responses(self._z, v, 500 + i)
            t0 = time.perf_counter()
            orc.rdm(x)
            total += time.perf_counter() - t0
        return total

    def time_replicates(self) -> float:
        """Wall seconds for cores x reps_per_worker replicates fanned over the pool."""
        jobs = [(w * self.reps_per_worker, self.reps_per_worker) for w in range(self.cores)]
        t0 = time.perf_counter()
        self.pool.map(_worker, jobs)
        return time.perf_counter() - t0

    def time_serial(self, reps: int = 8) -> float:
        """Seconds per replicate of the loop AS WRITTEN (visreps/evals.py:361-369): one Python process, replicates one
        after the other, BLAS threads = cores, SciPy ranking single-threaded."""
        torch.set_num_threads(self.cores)
        a, b, idx = _shared["a"], _shared["b"], _shared["idx"]
        t0 = time.perf_counter()
        orc.bootstrap_compare_with_indices(a, b, "spearman", idx[:reps])
        return (time.perf_counter() - t0) / reps

    def as_written(self, t_rdm: float, reps: int = 8) -> Dict[str, float]:
        """The serial reference job (BASELINE.md 3.4 'reference as written'), scaled from `reps` replicates."""
        per = self.time_serial(reps)
        n_pairs = len(self.layer_dims) * len(self.roi_dims)
        job = t_rdm + n_pairs * (self.n_bootstrap + 1) * per
        return {"value": n_pairs * self.n_bootstrap / job, "job_seconds": job, "s_per_replicate": per, "replicates_timed": reps,
                "cores": 1, "note": "serial Python loop as in evals.py:361-369 (SciPy ranking is single-threaded); RDM builds use all cores"}

    def step(self, t_rdm: float | None = None) -> Dict[str, float]:
        """One bounded sample -> estimated whole-job seconds and replicate-compares/s."""
        t_rdm = self.time_rdms() if t_rdm is None else t_rdm
        t_fan = self.time_replicates()
        n_pairs = len(self.layer_dims) * len(self.roi_dims)
        compares = n_pairs * (self.n_bootstrap + 1)          # + the point estimate of each pair
        per_compare = t_fan / (self.cores * self.reps_per_worker)
        job = t_rdm + compares * per_compare
        return {"job_seconds": job, "t_rdm": t_rdm, "t_fan": t_fan, "per_compare_s_per_core": t_fan / self.reps_per_worker,
                "value": n_pairs * self.n_bootstrap / job}

    def sample_description(self) -> str:
        return (f"{len(self.layer_dims) + len(self.roi_dims)} RDM builds (torch CPU, {self.cores} threads) + "
                f"{self.cores}x{self.reps_per_worker} Spearman replicates of one (layer, ROI) pair fanned over "
                f"{self.cores} processes, scaled linearly to {len(self.layer_dims) * len(self.roi_dims)} pairs x "
                f"{self.n_bootstrap} replicates")
