"""Test double of the two `visreps.utils` helpers the batched results sink takes from its host  --  TEST INFRASTRUCTURE.

    _compute_run_id   /root/reference/visreps/utils.py:304-309  (sha256 of the identity fields, first 12 hex digits)
    _init_db          /root/reference/visreps/utils.py:312-376  (opens results.db, creates the four tables)

The table definitions are not restated here: they are read from tests/golden/results_db.json, where
oracle/gen_golden_evals.py stored the `sqlite_master` rows of a database written by the unmodified `save_results`.
"""
from __future__ import annotations

import hashlib
import json
import sqlite3
import types
from pathlib import Path


def host_utils(golden: dict, db_path) -> types.ModuleType:
    fields = tuple(golden["identity_fields"])

    def _compute_run_id(cfg) -> str:
        identity = {f: cfg.get(f) for f in fields}
        identity["subject_idx"] = str(identity.get("subject_idx"))
        return hashlib.sha256(json.dumps(identity, sort_keys=True).encode()).hexdigest()[:12]

    def _init_db(path) -> sqlite3.Connection:
        conn = sqlite3.connect(str(path), timeout=10)
        for ddl in golden["schema"]:
            conn.execute(ddl.replace("CREATE TABLE ", "CREATE TABLE IF NOT EXISTS ", 1))
        conn.commit()
        return conn

    m = types.ModuleType("visreps.utils")
    m._compute_run_id, m._init_db, m._RESULTS_DB_PATH = _compute_run_id, _init_db, Path(db_path)
    return m


def dump_tables(db_path) -> dict:
    conn = sqlite3.connect(str(db_path))
    tables = {}
    for t in ("results", "run_configs", "layer_selection_scores", "bootstrap_distributions"):
        cur = conn.execute(f"SELECT * FROM {t}")
        cols = [c[0] for c in cur.description]
        rows = [dict(zip(cols, r)) for r in cur.fetchall()]
        for r in rows:
            r.pop("created_at", None)
        tables[t] = sorted(rows, key=lambda r: json.dumps(r, sort_keys=True, default=str))
    conn.close()
    return tables
