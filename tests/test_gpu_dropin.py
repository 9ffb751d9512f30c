"""The signature-level drop-ins for `_eval_rsa` / `_eval_rsa_nsd_synthetic` (visreps/evals.py:222, :442) against the
DataFrames the UNMODIFIED reference functions produced on the same synthetic job (tests/golden/eval_rsa.json, written by
oracle/gen_golden_evals.py), bound the way a user binds them: `visreps_b200.install()` on a `visreps.evals` module."""
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

import visreps_b200 as vb
from visreps_b200 import _lib, evals as vevals
from oracle import cases, results_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-5


class _OmegaConf:
    @staticmethod
    def merge(a, b):
        out = cases.EvalCfg(a)
        out.update(b)
        return out


@pytest.fixture
def host(monkeypatch):
    """A `visreps.evals`-shaped module: the reference's module-level names the loops reach, with a fake producer."""
    import pandas as pd
    saved = []
    ev = types.ModuleType("visreps.evals")
    ev.rprint = lambda *a, **k: None
    ev.pd = pd
    ev.OmegaConf = _OmegaConf
    ev.get_transform = lambda ds_stats="imgnet": None
    ev._make_loader = lambda stimuli, transform, batchsize, num_workers: stimuli
    ev.reconstruct_from_pcs = None
    ev.save_results = lambda df, cfg: saved.append((df.to_dict("records"), dict(cfg)))
    ev.mutils = types.SimpleNamespace(extract_single_layer=cases.fake_extract_single_layer,
                                      load_model=lambda cfg, dev, verbose=False: cfg["_model"],
                                      configure_feature_extractor=lambda cfg, m, verbose=False: m)
    ev._eval_rsa = ev._eval_rsa_nsd_synthetic = ev.compute_rdm = ev.compute_rdm_correlation = None
    ev.saved = saved
    for name in ("visreps", "visreps.analysis", "visreps.analysis.rsa", "visreps.analysis.alignment"):
        monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    monkeypatch.setitem(sys.modules, "visreps.evals", ev)
    vb.install()
    yield ev
    vb.uninstall()


def _check_rows(got_rows, want_rows, bootstrap):
    assert len(got_rows) == len(want_rows)
    for g, w in zip(got_rows, want_rows):
        assert list(g.keys()) == list(w.keys())
        assert g["layer"] == w["layer"] and g["compare_method"] == w["compare_method"] and g["analysis"] == "rsa"
        assert isinstance(g["score"], float) and abs(g["score"] - w["score"]) <= TOL
        assert [e["layer"] for e in g["layer_selection_scores"]] == [e["layer"] for e in w["layer_selection_scores"]]
        for a, b in zip(g["layer_selection_scores"], w["layer_selection_scores"]):
            assert abs(a["score"] - b["score"]) <= TOL
        if bootstrap:
            assert isinstance(g["bootstrap_scores"], list) and len(g["bootstrap_scores"]) == len(w["bootstrap_scores"])
            np.testing.assert_allclose(g["bootstrap_scores"], w["bootstrap_scores"], atol=TOL, rtol=0)
            assert abs(g["ci_low"] - w["ci_low"]) <= TOL and abs(g["ci_high"] - w["ci_high"]) <= TOL
        else:
            assert g["ci_low"] is None and g["ci_high"] is None and "bootstrap_scores" not in g


@pytest.mark.parametrize("job", ["spearman_boot", "pearson_boot", "kendall_point", "spearman_all_train"])
def test_eval_rsa_drop_in_matches_the_unmodified_reference(job, host, golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "eval_rsa.json")))["jobs"][job]
    cfg, model, acts, ids, all_data, subjects, regions = cases.eval_rsa_job(**gold["kwargs"])
    assert host._eval_rsa is vevals._eval_rsa                                      # bound by install()
    l0 = _lib.launch_count()
    df = host._eval_rsa(cfg, model, dict(acts), ids, all_data, subjects, regions, torch.device("cuda", 0), False)
    launches = _lib.launch_count() - l0
    assert launches > 0
    _check_rows(df.to_dict("records"), gold["rows"], gold["kwargs"].get("bootstrap", True))
    # save_results was called once per (region, subject), in the reference's order, with the merged cfg
    assert [(c["region"], c["subject_idx"]) for _, c in host.saved] == [(s["region"], s["subject_idx"]) for s in gold["saved"]]
    for (rows, _), s in zip(host.saved, gold["saved"]):
        _check_rows(rows, s["rows"], gold["kwargs"].get("bootstrap", True))


def test_eval_rsa_with_device_resident_producer_and_batched_sink(host, golden_dir, tmp_path):
    """Activations that never leave the GPU (visreps_b200.producer's hand-off) give the same DataFrame, phase 1 shares
    the model-side RDMs between the regions of a subject, and the batched sink stores what the reference stores."""
    dev = torch.device("cuda", 0)
    gold = json.load(open(os.path.join(golden_dir, "eval_rsa.json")))["jobs"]["spearman_boot"]
    cfg, model, acts, ids, all_data, subjects, regions = cases.eval_rsa_job(**gold["kwargs"])
    cfg.update(cases.results_cfg())
    acts_dev = {k: v.to(dev) for k, v in acts.items()}
    model_dev = {"exact": {k: v.to(dev) for k, v in model["exact"].items()}, "row": model["row"]}
    gold_db = json.load(open(os.path.join(golden_dir, "results_db.json")))
    utils = results_oracle.host_utils(gold_db, tmp_path / "results.db")
    from visreps_b200 import results_sink
    vevals.RESULTS_SINK = lambda pairs: results_sink.save_results_batch(pairs, utils=utils)
    try:
        _lib.profile_start()
        df = host._eval_rsa(cfg, model_dev, acts_dev, ids, all_data, subjects, regions, dev, False)
        prof = _lib.profile_stop()
    finally:
        vevals.RESULTS_SINK = None
    _check_rows(df.to_dict("records"), gold["rows"], True)
    assert host.saved == []                                                       # the sink replaced the per-pair calls
    got = results_oracle.dump_tables(tmp_path / "results.db")
    assert [r["run_id"] for r in got["results"]] == [r["run_id"] for r in gold_db["tables"]["results"]]
    for g, w in zip(got["results"], gold_db["tables"]["results"]):
        assert g["layer"] == w["layer"] and abs(g["score"] - w["score"]) <= TOL and abs(g["ci_low"] - w["ci_low"]) <= TOL
        assert {k: g[k] for k in ("seed", "epoch", "region", "subject_idx", "neural_dataset", "cfg_id", "model_name", "pca_k")} == \
               {k: w[k] for k in ("seed", "epoch", "region", "subject_idx", "neural_dataset", "cfg_id", "model_name", "pca_k")}
    assert len(got["layer_selection_scores"]) == len(gold_db["tables"]["layer_selection_scores"])
    # phase 1: 3 layers x 2 subjects = 6 model-side RDM builds (+ 4 neural), not 3 x 4 pairs; phase 2: 1-2 layers + 4 neural
    syrk = prof["rdm_syrk"]["launches"] + prof["rdm_syrk_pair"]["launches"]
    n_layers_scored = len({r["layer"] for r in gold["rows"]})
    assert syrk == (3 * 2 + 4) + (n_layers_scored + 4), (syrk, prof)


def test_eval_rsa_nsd_synthetic_drop_in(host, golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "eval_rsa.json")))["nsd_synthetic"]
    cfg, model, acts, ids, all_data, subjects, regions = cases.eval_rsa_job(seed=91, n_shared=44, n_private=10, bootstrap=True, n_bootstrap=10)
    layers = {r: {int(s): l for s, l in d.items()} for r, d in gold["layers"].items()}
    test_ids = all_data["shared_test_ids"]
    cfg["_model"] = model
    host._lookup_nsd_best_layers = lambda cfg, subjects, regions: layers
    host.load_nsd_synthetic_test_data = lambda cfg, subjects, regions: {
        "test_ids": test_ids, "stimuli": {s: all_data["stimuli"][s] for s in test_ids},
        "neural": {r: {s: all_data["neural"][r][s]["test"] for s in subjects} for r in regions}}
    df = host._eval_rsa_nsd_synthetic(cfg, subjects, regions, torch.device("cuda", 0), False)
    _check_rows(df.to_dict("records"), gold["rows"], True)
    assert len(host.saved) == gold["n_saved"]
    assert all(r["layer_selection_scores"] == [] for r in df.to_dict("records"))


def test_compute_rsa_keeps_the_observable_call_sequence_when_the_comparison_is_wrapped(monkeypatch):
    """tests/test_rsa_bootstrap.py:519-546 wraps compute_rdm_correlation and counts the sub-RDM sizes it sees."""
    from visreps_b200 import rsa as vrsa
    train, test = cases.alignment_case()
    sel, ev = vb.AlignmentData(**train), vb.AlignmentData(**test)
    plain = vb.compute_rsa({"compare_method": "spearman"}, sel, ev, n_select=100, bootstrap=True, n_bootstrap=5, seed=42)[0]
    sizes = []
    original = vrsa.compute_rdm_correlation

    def tracking(rdm1, rdm2, **kw):
        sizes.append(int(rdm1.size(0)))
        return original(rdm1, rdm2, **kw)

    monkeypatch.setattr(vrsa, "compute_rdm_correlation", tracking)
    seen = vb.compute_rsa({"compare_method": "spearman"}, sel, ev, n_select=100, bootstrap=True, n_bootstrap=5, seed=42)[0]
    n_test = test["neural"].shape[0]
    assert sizes.count(int(n_test * 0.9)) == 5 and sizes.count(n_test) == 1
    assert seen["layer"] == plain["layer"] and abs(seen["score"] - plain["score"]) <= 1e-12
    np.testing.assert_allclose(seen["bootstrap_scores"], plain["bootstrap_scores"], atol=1e-12, rtol=0)
