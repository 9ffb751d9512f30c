"""Golden results of the UNMODIFIED reference `_eval_rsa` / `_eval_rsa_nsd_synthetic` (test infrastructure).

    python oracle/gen_golden_evals.py          # build container only: needs /root/reference

`visreps/evals.py` imports the whole application (omegaconf, h5py, xarray, himalaya, wandb ...), none of which is on
the numeric path of these two functions.  The script pre-seeds stubs for exactly those imports -- `omegaconf`,
`visreps.utils`, `visreps.models.utils`, `visreps.dataloaders.*`, `visreps.analysis.encoding_score`,
`visreps.analysis.reconstruct_from_pcs` -- and imports evals.py, alignment.py and rsa.py AS THEY ARE.  The fake
producer (`oracle.cases.fake_extract_single_layer`) and the synthetic job (`oracle.cases.eval_rsa_job`) are the ones the
tests rebuild, so tests/golden/eval_rsa.json pins the drop-ins in visreps_b200/evals.py to the reference's own loops.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REFERENCE = os.environ.get("VISREPS_REFERENCE", "/root/reference")

SAVED = []


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference_evals(cases):
    from rich.console import Console

    class OmegaConf:
        @staticmethod
        def merge(a, b):
            out = cases.EvalCfg(a)
            out.update(b)
            return out

        load = create = to_container = staticmethod(lambda *a, **k: (_ for _ in ()).throw(RuntimeError("not on this path")))

    _stub("omegaconf", OmegaConf=OmegaConf, ListConfig=list, DictConfig=dict)
    pkg = _stub("visreps"); pkg.__path__ = [os.path.join(REFERENCE, "visreps")]
    ana = _stub("visreps.analysis"); ana.__path__ = [os.path.join(REFERENCE, "visreps", "analysis")]
    _stub("visreps.utils", console=Console(quiet=True), rprint=lambda *a, **k: None,
          save_results=lambda df, cfg: SAVED.append((df.to_dict("records"), dict(cfg))),
          _compute_run_id=lambda cfg: "run", get_seed_letter=lambda s: "a")
    _stub("visreps.models"); sys.modules["visreps.models"].__path__ = []
    _stub("visreps.models.utils", extract_single_layer=cases.fake_extract_single_layer, TORCHVISION_RETURN_NODES={},
          load_model=lambda cfg, dev, verbose=False: cfg["_model"], configure_feature_extractor=lambda cfg, m, verbose=False: m)
    _stub("visreps.dataloaders"); sys.modules["visreps.dataloaders"].__path__ = []
    never = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("not on this path"))
    _stub("visreps.dataloaders.neural", get_neural_loader=never, load_all_nsd_data=never, load_all_tvsd_data=never,
          load_nsd_synthetic_test_data=lambda cfg, subjects, regions: cfg["_synthetic"],
          _make_loader=lambda stimuli, transform, batchsize, num_workers: stimuli)
    _stub("visreps.dataloaders.obj_cls", get_transform=lambda ds_stats="imgnet": None)
    _stub("visreps.analysis.encoding_score", compute_encoding_score=never)
    _stub("visreps.analysis.reconstruct_from_pcs", reconstruct_from_pcs=never)
    import importlib
    return importlib.import_module("visreps.evals")


def _clean(rows):
    out = []
    for r in rows:
        r = dict(r)
        for k, v in list(r.items()):
            if isinstance(v, float) and np.isnan(v):
                r[k] = None
        out.append(r)
    return out


def main():
    import torch
    from oracle import cases
    ev = import_reference_evals(cases)
    dev = torch.device("cpu")
    golden = {"reference": REFERENCE, "jobs": {}}
    jobs = {
        "spearman_boot": dict(method="spearman", bootstrap=True, n_bootstrap=12),
        "pearson_boot": dict(method="pearson", bootstrap=True, n_bootstrap=6),
        "kendall_point": dict(method="kendall", bootstrap=False, n_shared=40, n_private=70, n_select=50),
        "spearman_all_train": dict(method="spearman", bootstrap=False, n_select=100000),
    }
    for name, kw in jobs.items():
        cfg, model, acts, ids, all_data, subjects, regions = cases.eval_rsa_job(**kw)
        SAVED.clear()
        df = ev._eval_rsa(cfg, model, dict(acts), ids, all_data, subjects, regions, dev, False)
        golden["jobs"][name] = {"kwargs": kw, "rows": _clean(df.to_dict("records")),
                                "saved": [{"rows": _clean(rows), "subject_idx": c["subject_idx"], "region": c["region"]} for rows, c in SAVED]}
        print(name, [(r["layer"], round(r["score"], 4)) for r in golden["jobs"][name]["rows"]])
    # NSD-synthetic: best layers looked up (stubbed), 220-stimulus-like test set
    cfg, model, acts, ids, all_data, subjects, regions = cases.eval_rsa_job(seed=91, n_shared=44, n_private=10, bootstrap=True, n_bootstrap=10)
    layers = {"V1": {0: "conv1_post", 1: "fc2_post"}, "ventral": {0: "fc2_post", 1: "fc2_post"}}
    ev._lookup_nsd_best_layers = lambda cfg, subjects, regions: layers
    test_ids = all_data["shared_test_ids"]
    cfg["_model"] = model
    cfg["_synthetic"] = {"test_ids": test_ids, "stimuli": {s: all_data["stimuli"][s] for s in test_ids},
                         "neural": {r: {s: all_data["neural"][r][s]["test"] for s in subjects} for r in regions}}
    SAVED.clear()
    df = ev._eval_rsa_nsd_synthetic(cfg, subjects, regions, dev, False)
    golden["nsd_synthetic"] = {"layers": {r: {str(s): l for s, l in d.items()} for r, d in layers.items()},
                               "rows": _clean(df.to_dict("records")), "n_saved": len(SAVED)}
    path = os.path.join(ROOT, "tests", "golden", "eval_rsa.json")
    with open(path, "w") as f:
        json.dump(golden, f)
    print("wrote", path, os.path.getsize(path), "bytes")


def results_db_golden():
    """Run the UNMODIFIED `save_results` (visreps/utils.py:381-458) on the rows of one golden job and dump the database:
    schema (as data), run ids and every table.  visreps/utils.py is imported as it is, with stubs for omegaconf, dotenv
    and wandb (none of them on the path of save_results beyond OmegaConf.to_container)."""
    import importlib
    import sqlite3
    import tempfile
    from pathlib import Path
    import pandas as pd
    from oracle import cases
    for name in [m for m in sys.modules if m == "visreps" or m.startswith("visreps.")]:
        del sys.modules[name]

    class OmegaConf:
        to_container = staticmethod(lambda cfg, resolve=True: dict(cfg))
        merge = staticmethod(lambda a, b: cases.EvalCfg({**a, **b}))

    _stub("omegaconf", OmegaConf=OmegaConf, ListConfig=list, DictConfig=dict)
    _stub("dotenv", load_dotenv=lambda *a, **k: None)
    _stub("wandb")
    pkg = _stub("visreps"); pkg.__path__ = [os.path.join(REFERENCE, "visreps")]
    utils = importlib.import_module("visreps.utils")
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "eval_rsa.json")))
    job = golden["jobs"]["spearman_boot"]
    base = cases.results_cfg()
    with tempfile.TemporaryDirectory() as tmp:
        utils._RESULTS_DB_PATH = Path(tmp) / "results.db"
        run_ids = []
        for item in job["saved"]:
            cfg = cases.EvalCfg({**base, "subject_idx": item["subject_idx"], "region": item["region"]})
            utils.save_results(pd.DataFrame(item["rows"]), cfg)
            run_ids.append(utils._compute_run_id(cfg))
        conn = sqlite3.connect(str(utils._RESULTS_DB_PATH))
        schema = [r[0] for r in conn.execute("SELECT sql FROM sqlite_master WHERE type='table' ORDER BY name") if r[0]]
        tables = {}
        for t in ("results", "run_configs", "layer_selection_scores", "bootstrap_distributions"):
            cur = conn.execute(f"SELECT * FROM {t}")
            cols = [c[0] for c in cur.description]
            rows = [dict(zip(cols, r)) for r in cur.fetchall()]
            for r in rows:
                r.pop("created_at", None)
            tables[t] = sorted(rows, key=lambda r: json.dumps(r, sort_keys=True, default=str))
        conn.close()
    out = {"schema": schema, "identity_fields": list(utils._IDENTITY_FIELDS), "run_ids": run_ids, "tables": tables}
    path = os.path.join(ROOT, "tests", "golden", "results_db.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes;", {t: len(v) for t, v in tables.items()})


if __name__ == "__main__":
    main()
    results_db_golden()
