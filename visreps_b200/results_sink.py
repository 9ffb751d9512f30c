"""Batched results sink: all (region, subject) rows of one eval job in ONE SQLite transaction.

The reference calls `save_results(df, cfg)` once per pair (visreps/evals.py:390-392, :546-548); each call opens the
database, creates the tables if needed, inserts a handful of rows and commits (visreps/utils.py:381-458) -- 48 WAL
commits per NSD job.  `save_results_batch` writes the same rows into the same tables (`run_configs`, `results`,
`layer_selection_scores`, `bootstrap_distributions`; same columns, same INSERT OR REPLACE keys, same run_id) through
one connection and one commit.  The schema and the run-id hash stay the reference's: `_init_db` and
`_compute_run_id` are taken from the host module (`visreps.utils`).

Bootstrap distributions are stored as JSON text exactly like the reference by default, so existing readers
(`json.loads(scores)`) keep working; `blob=True` stores the little-endian float64 vector instead (8 B per
replicate instead of ~20) for consumers that opt in with `decode_scores`.
"""
from __future__ import annotations

import json
import math
import sys
from typing import Iterable, Optional, Sequence, Tuple

import numpy as np


def _utils_module(utils=None):
    if utils is not None:
        return utils
    mod = sys.modules.get("visreps.utils")
    if mod is None:
        raise RuntimeError("visreps.utils is not imported; pass utils=<module providing _init_db, _compute_run_id, _RESULTS_DB_PATH>")
    return mod


def _to_container(cfg, utils):
    oc = getattr(utils, "OmegaConf", None)
    if oc is not None:
        try:
            return oc.to_container(cfg, resolve=True)
        except Exception:
            pass
    return {k: cfg[k] for k in cfg.keys()}


def _float_or_none(row, col):
    v = row.get(col) if hasattr(row, "get") else None
    if v is None:
        return None
    try:
        v = float(v)
    except (TypeError, ValueError):
        return None
    return None if math.isnan(v) else v


def decode_scores(stored) -> list:
    """Bootstrap scores from either encoding of the `scores` column."""
    if isinstance(stored, (bytes, memoryview)):
        return np.frombuffer(bytes(stored), dtype="<f8").tolist()
    return json.loads(stored)


def save_results_batch(pairs: Iterable[Tuple[object, object]], utils=None, db_path=None, blob: bool = False) -> str:
    """pairs: (DataFrame-or-list-of-row-dicts, cfg) per (region, subject), what the reference passes to save_results."""
    utils = _utils_module(utils)
    path = db_path if db_path is not None else utils._RESULTS_DB_PATH
    conn = utils._init_db(path)
    n_rows = 0
    try:
        with conn:                                             # one transaction
            for df, cfg in pairs:
                rows = df.to_dict("records") if hasattr(df, "to_dict") else list(df)
                run_id = utils._compute_run_id(cfg)
                conn.execute("INSERT OR REPLACE INTO run_configs (run_id, config_json) VALUES (?, ?)",
                             (run_id, json.dumps(_to_container(cfg, utils), default=str)))
                for row in rows:
                    method = row.get("compare_method", cfg.get("compare_method", "spearman"))
                    score = _float_or_none(row, "score")
                    if score is not None:
                        conn.execute(
                            """INSERT OR REPLACE INTO results
                               (run_id, compare_method, layer, score, ci_low, ci_high,
                                analysis, seed, epoch, region, subject_idx,
                                neural_dataset, cfg_id, pca_labels, pca_n_classes, pca_labels_folder,
                                model_name, checkpoint_dir, reconstruct_from_pcs, pca_k)
                               VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?)""",
                            (run_id, method, row.get("layer"), score, _float_or_none(row, "ci_low"), _float_or_none(row, "ci_high"),
                             row.get("analysis", cfg.get("analysis")), int(cfg.get("seed")), int(cfg.get("epoch", 0)),
                             cfg.get("region"), str(cfg.get("subject_idx")), cfg.get("neural_dataset"), cfg.get("cfg_id"),
                             bool(cfg.get("pca_labels")), cfg.get("pca_n_classes"), cfg.get("pca_labels_folder"),
                             cfg.get("model_name"), cfg.get("checkpoint_dir"), bool(cfg.get("reconstruct_from_pcs", False)),
                             cfg.get("pca_k", 1)))
                        n_rows += 1
                    sel = row.get("layer_selection_scores")
                    sel = sel if isinstance(sel, (list, tuple)) else []
                    conn.executemany(
                        "INSERT OR REPLACE INTO layer_selection_scores (run_id, compare_method, layer, score) VALUES (?, ?, ?, ?)",
                        [(run_id, method, e["layer"], float(e["score"])) for e in sel])
                    bs = row.get("bootstrap_scores")
                    if isinstance(bs, (list, tuple, np.ndarray)):
                        payload = np.asarray(bs, dtype="<f8").tobytes() if blob else json.dumps(list(map(float, bs)))
                        conn.execute("INSERT OR REPLACE INTO bootstrap_distributions (run_id, compare_method, scores) VALUES (?, ?, ?)",
                                     (run_id, method, payload))
    finally:
        conn.close()
    say = getattr(utils, "rprint", None)
    if say is not None:
        say(f"Saved {n_rows} results to {path} in one transaction", style="success")
    return str(path)
