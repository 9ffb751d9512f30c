"""Host logic of the drop-in layer (no GPU): install() bindings and warnings, the batched results sink against a
database written by the unmodified reference `save_results`, the producer shim's bookkeeping."""
import json
import os
import sqlite3
import sys
import types
import warnings

import numpy as np
import pytest
import torch

import visreps_b200 as vb
from visreps_b200 import evals as vevals, producer, results_sink
vinstall = sys.modules["visreps_b200.install"]
from oracle import cases, results_oracle


@pytest.fixture
def fake_tree(monkeypatch):
    sentinel = object()
    mods = {}
    spec = {"visreps": (), "visreps.analysis": (), "visreps.models": (),
            "visreps.analysis.rsa": ("compute_rdm", "compute_rdm_correlation", "compute_rsa", "_rank", "_concept_average_exact"),
            "visreps.evals": ("compute_rdm", "compute_rdm_correlation", "_concept_average_exact", "_eval_rsa", "_eval_rsa_nsd_synthetic", "eval"),
            "visreps.analysis.alignment": ("compute_rsa",),
            "visreps.models.utils": ("get_activations", "extract_single_layer", "load_model")}
    for name, syms in spec.items():
        m = types.ModuleType(name)
        for s in syms:
            setattr(m, s, sentinel)
        mods[name] = m
        monkeypatch.setitem(sys.modules, name, m)
    yield mods, sentinel
    vb.uninstall()


def test_install_binds_the_eval_loops_and_optionally_the_producer(fake_tree):
    mods, sentinel = fake_tree
    done = vb.install()
    assert ("visreps.evals", "_eval_rsa") in done and ("visreps.evals", "_eval_rsa_nsd_synthetic") in done
    assert mods["visreps.evals"]._eval_rsa is vevals._eval_rsa
    assert mods["visreps.evals"]._eval_rsa_nsd_synthetic is vevals._eval_rsa_nsd_synthetic
    assert mods["visreps.evals"].eval is sentinel                           # the orchestration stays the reference's
    assert mods["visreps.models.utils"].get_activations is sentinel          # producer untouched unless asked
    vb.uninstall()
    assert mods["visreps.evals"]._eval_rsa is sentinel
    done = vb.install(device_producer=True, batched_sink=True)
    assert mods["visreps.models.utils"].get_activations is producer.get_activations
    assert mods["visreps.models.utils"].extract_single_layer is producer.extract_single_layer
    assert mods["visreps.models.utils"].load_model is sentinel
    assert vevals.RESULTS_SINK is results_sink.save_results_batch
    vb.uninstall()
    assert mods["visreps.models.utils"].extract_single_layer is sentinel and vevals.RESULTS_SINK is None


def test_install_reports_modules_it_cannot_import(monkeypatch):
    for name in list(sys.modules):
        if name == "visreps" or name.startswith("visreps."):
            monkeypatch.delitem(sys.modules, name)
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        done = vb.install()
    assert done == []
    msgs = [str(w.message) for w in caught if issubclass(w.category, vinstall.InstallWarning)]
    assert len(msgs) == 3 and any("visreps.evals" in m and "_eval_rsa" in m for m in msgs)
    with pytest.raises(ImportError):
        vb.install(strict=True)


def test_eval_signatures_match_the_reference():
    import inspect
    assert list(inspect.signature(vevals._eval_rsa).parameters)[:9] == ["cfg", "model", "acts", "ids", "all_data", "subjects", "regions", "dev", "verbose"]
    assert list(inspect.signature(vevals._eval_rsa_nsd_synthetic).parameters)[:5] == ["cfg", "subjects", "regions", "dev", "verbose"]
    assert list(inspect.signature(producer.get_activations).parameters)[:3] == ["model", "dataloader", "device"]
    assert list(inspect.signature(producer.extract_single_layer).parameters)[:5] == ["model", "dataloader", "device", "layer_name", "stimulus_ids"]
    with pytest.raises(RuntimeError):                                        # no host module: loud, not a silent fallback
        saved = sys.modules.pop("visreps.evals", None)
        try:
            vevals._eval_rsa({}, None, {}, [], {}, [], [], None, False)
        finally:
            if saved is not None:
                sys.modules["visreps.evals"] = saved


def test_batched_sink_writes_what_the_reference_save_results_writes(golden_dir, tmp_path):
    import pandas as pd
    gold_db = json.load(open(os.path.join(golden_dir, "results_db.json")))
    job = json.load(open(os.path.join(golden_dir, "eval_rsa.json")))["jobs"]["spearman_boot"]
    utils = results_oracle.host_utils(gold_db, tmp_path / "results.db")
    base = cases.results_cfg()
    pairs = [(pd.DataFrame(item["rows"]), cases.EvalCfg({**base, "subject_idx": item["subject_idx"], "region": item["region"]}))
             for item in job["saved"]]
    assert [utils._compute_run_id(c) for _, c in pairs] == gold_db["run_ids"]
    path = results_sink.save_results_batch(pairs, utils=utils)
    got = results_oracle.dump_tables(path)
    for t in ("results", "layer_selection_scores", "bootstrap_distributions"):
        assert got[t] == gold_db["tables"][t], t
    assert [json.loads(r["config_json"]) for r in got["run_configs"]] == [json.loads(r["config_json"]) for r in gold_db["tables"]["run_configs"]]
    # re-running replaces, never duplicates (INSERT OR REPLACE keys of the reference)
    results_sink.save_results_batch(pairs, utils=utils)
    assert results_oracle.dump_tables(path)["results"] == gold_db["tables"]["results"]
    # rows given as plain dicts, float64 blob encoding on request
    blob_path = results_sink.save_results_batch([(item["rows"], c) for item, (_, c) in zip(job["saved"], pairs)], utils=utils,
                                                db_path=tmp_path / "blob.db", blob=True)
    conn = sqlite3.connect(blob_path)
    stored = conn.execute("SELECT scores FROM bootstrap_distributions ORDER BY run_id").fetchall()
    conn.close()
    want = sorted((json.loads(r["scores"]) for r in gold_db["tables"]["bootstrap_distributions"]), key=lambda v: v[0])
    assert sorted((results_sink.decode_scores(s[0]) for s in stored), key=lambda v: v[0]) == want
    # a row without a score is skipped like the reference does (utils.py:401-402), NaN CI bounds become NULL
    odd = [{"layer": "l", "compare_method": "spearman", "score": float("nan"), "analysis": "rsa"},
           {"layer": "m", "compare_method": "spearman", "score": 0.5, "ci_low": float("nan"), "ci_high": None, "analysis": "rsa"}]
    results_sink.save_results_batch([(odd, pairs[0][1])], utils=utils, db_path=tmp_path / "odd.db")
    rows = results_oracle.dump_tables(tmp_path / "odd.db")["results"]
    assert len(rows) == 1 and rows[0]["layer"] == "m" and rows[0]["ci_low"] is None and rows[0]["ci_high"] is None


class _Loader:
    def __init__(self, batches):
        self.batches = batches
        self.dataset = [None] * sum(len(k) for _, k in batches)

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)


class _Net(torch.nn.Module):
    def forward(self, x):
        return {"conv": x.reshape(x.size(0), 2, -1) * 2.0, "fc": x[:, :3] + 1.0}


def test_producer_shim_keeps_order_layers_and_dtype():
    """CPU stand-in for the device: same rows, same order, same SRP arithmetic as the reference loop, no list of
    per-batch host copies."""
    import scipy.sparse as sp
    rs = np.random.RandomState(0)
    imgs = torch.from_numpy(rs.standard_normal((10, 8)).astype(np.float32))
    batches = [(imgs[:4], ["a", "b", "c", "d"]), (imgs[4:8], ["e", "f", "g", "h"]), (imgs[8:], ["i", "j"])]
    mats = {}

    def factory(D, k, density, seed, cache_dir):
        m = sp.random(k, D, density=0.5, random_state=D, format="csr", dtype=np.float32)
        mats[D] = m
        return types.SimpleNamespace(components_=m)

    dev = torch.device("cpu")
    acts, ids = producer.get_activations(_Net(), _Loader(batches), dev, srp_factory=factory)
    assert ids == list("abcdefghij") and set(acts) == {"conv", "fc"}
    want_conv = (imgs * 2.0).numpy() @ mats[8].toarray().T
    np.testing.assert_allclose(acts["conv"].numpy(), want_conv, rtol=1e-6, atol=1e-6)
    assert acts["fc"].shape == (10, 3) and acts["fc"].dtype == torch.float32
    exact, order = producer.extract_single_layer(_Net(), _Loader(batches), dev, "conv", ["j", "zz", "a", "e"])
    assert order == ["j", "a", "e"] and torch.equal(exact, (imgs * 2.0)[[9, 0, 4]])
    exact_all, order_all = producer.extract_single_layer(_Net(), _Loader(batches), dev, "fc")
    assert order_all == list("abcdefghij") and torch.equal(exact_all, imgs[:, :3] + 1.0)
    assert producer.get_activations(_Net(), _Loader([]), dev, srp_factory=factory) == ({}, [])


def test_align_stimulus_level_matches_alignment_py():
    """alignment.py:23-39: rows whose id has a target, in `keys` order; ids compared as strings; empty -> empty."""
    from visreps_b200.alignment import _align_stimulus_level
    cfg, model, acts, ids, all_data, subjects, regions = cases.eval_rsa_job()
    targets = all_data["neural"]["V1"][1]["train"]
    got_acts, neural, matched = _align_stimulus_level(acts, targets, ids)
    want_idx = [i for i, k in enumerate(ids) if k in targets]
    assert matched == [ids[i] for i in want_idx] and len(matched) == 150
    assert torch.equal(neural, torch.as_tensor(np.stack([targets[k] for k in matched]), dtype=torch.float32))
    assert all(torch.equal(got_acts[l], acts[l][want_idx]) for l in acts) and neural.dtype == torch.float32
    int_keys = [7, 8, 9]
    a2, n2, m2 = _align_stimulus_level({"l": torch.arange(6.).reshape(3, 2)}, {"8": np.ones(4), "9": np.zeros(4)}, int_keys)
    assert m2 == ["8", "9"] and torch.equal(a2["l"], torch.tensor([[2., 3.], [4., 5.]])) and n2.shape == (2, 4)
    a3, n3, m3 = _align_stimulus_level({"l": torch.zeros(3, 2)}, {}, int_keys)
    assert m3 == [] and a3["l"].shape == (0, 2) and n3.numel() == 0 and n3.dtype == torch.float32


def test_compute_traintest_alignment_dispatch(monkeypatch):
    """alignment.py:74-114: rsa -> compute_rsa with the cfg's n_select / bootstrap / n_bootstrap; the error cases."""
    from visreps_b200 import alignment as val
    seen = {}

    def fake(cfg, train, test, **kw):
        seen.update(kw)
        return [{"ok": True}]

    monkeypatch.setattr(val, "compute_rsa", fake)
    fn = object()
    assert val.compute_traintest_alignment({"analysis": "RSA", "n_select": 7, "bootstrap": False, "n_bootstrap": 3}, "tr", "te",
                                           verbose=True, re_extract_fn=fn) == [{"ok": True}]
    assert seen == {"n_select": 7, "bootstrap": False, "n_bootstrap": 3, "verbose": True, "re_extract_fn": fn}
    val.compute_traintest_alignment({}, "tr", "te")
    assert seen == {"n_select": None, "bootstrap": True, "n_bootstrap": 1000, "verbose": False, "re_extract_fn": None}
    with pytest.raises(ValueError):
        val.compute_traintest_alignment({"analysis": "encoding_score", "neural_dataset": "THINGS-behavior"}, "tr", "te")
    with pytest.raises(NotImplementedError):
        val.compute_traintest_alignment({"analysis": "encoding_score", "neural_dataset": "nsd"}, "tr", "te")
    with pytest.raises(ValueError):
        val.compute_traintest_alignment({"analysis": "cka"}, "tr", "te")


def test_concept_average_exact_matches_the_oracle():
    """rsa.py:284-305 restated in oracle.rsa_oracle.concept_average: ragged concepts, unseen images, a concept with no
    image at all (zeros), integer image ids, fp16 input; gathered in bounded blocks."""
    from oracle import rsa_oracle as orc
    rs = np.random.RandomState(4)
    n_img, d = 57, 300
    raw = torch.from_numpy(rs.standard_normal((n_img, d)).astype(np.float32))
    raw_ids = [f"img{k}" for k in rs.permutation(n_img)]
    concepts = [f"c{k}" for k in range(9)]
    members, pool = {}, list(raw_ids)
    for k, c in enumerate(concepts):
        members[c] = [pool.pop() for _ in range(1 + (k * 3) % 7)] + ([f"missing{k}"] if k % 2 else [])
    members["c4"] = ["nowhere1", "nowhere2"]
    data = types.SimpleNamespace(stimulus_ids=concepts, concept_image_ids=members)
    got = vb._concept_average_exact(raw, raw_ids, data)
    want = orc.concept_average(raw, raw_ids, data)
    assert got.shape == (9, d) and got.dtype == raw.dtype
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=0, atol=1e-6)
    assert float(got[4].abs().max()) == 0.0
    half = vb._concept_average_exact(raw.half(), raw_ids, data)
    assert half.dtype == torch.float16
    np.testing.assert_allclose(half.float().numpy(), orc.concept_average(raw.half(), raw_ids, data).float().numpy(), rtol=0, atol=2e-3)
