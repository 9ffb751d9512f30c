"""Device versions of the scoring blocks that `visreps/evals.py` inlines around the RSA numerics.

The reference evaluates NSD/TVSD in three places that all run the same loops over (region, subject):

  * phase 1, layer selection (evals.py:249-287): for every pair, up to `n_select` training stimuli drawn by a
    fresh `RandomState(42)`, one Pearson RDM per layer, one comparison per layer, first maximum wins;
  * the scoring block of `_eval_rsa` (evals.py:333-398) and of `_eval_rsa_nsd_synthetic` (evals.py:489-546):
    per pair a neural RDM, the point estimate against the pair's best layer and, optionally, 1000 90 %-subsample
    replicates from a `RandomState(42)` that is re-created for every pair -- so all pairs see the same index sets.

Here each block is one grouped device job (`rsa_grouped`): all layers x the pair's neural matrix for the
selection, all distinct best layers x all pairs for the scoring, with the rankings shared across pairs.  The
result dictionaries have the reference's keys and value types, in the reference's (region, subject) order.
Model loading, feature extraction, the neural data loaders and `save_results` stay where they are.
"""
from __future__ import annotations

from typing import Dict, Hashable, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import rsa

SELECTION_SEED = 42      # evals.py:259 and :356 hard-code RandomState(42)


def selection_indices(n_train: int, n_select: Optional[int]) -> np.ndarray:
    """evals.py:257-262: `RandomState(42).choice(n_train, n_select, replace=False)` when the pair has more
    training stimuli than `n_select`, else all of them in order."""
    if n_select is not None and n_select < n_train:
        return np.random.RandomState(SELECTION_SEED).choice(n_train, size=n_select, replace=False)
    return np.arange(n_train)


def select_layer(train_acts: Mapping[str, torch.Tensor], train_neural: torch.Tensor, method: str = "spearman",
                 n_select: Optional[int] = 1000) -> Tuple[Optional[str], List[Dict]]:
    """Phase-1 layer selection for one (region, subject) pair (evals.py:251-276).

    train_acts: {layer: (n_train, ...)} activations aligned with train_neural (n_train, V).
    Returns (best_layer, [{"layer": name, "score": float}, ...]) with the reference's tie rule: the first
    layer whose score is strictly greater than every earlier one (a NaN score never wins)."""
    n_train = int(train_neural.size(0))
    sel_idx = selection_indices(n_train, n_select)
    idx_t = torch.from_numpy(sel_idx)
    names = list(train_acts.keys())
    feats = [rsa._flat(train_acts[k][idx_t.to(train_acts[k].device)]) for k in names]
    res = rsa.rsa_grouped(feats, [train_neural[idx_t.to(train_neural.device)]], method, bootstrap=False)
    scores, best_layer, best_score = [], None, -float("inf")
    for name, s in zip(names, res["point"][:, 0]):
        s = float(s)
        scores.append({"layer": name, "score": s})
        if s > best_score:
            best_score, best_layer = s, name
    return best_layer, scores


def score_pairs(model_features: Mapping[str, torch.Tensor], neural_test: Mapping[Hashable, torch.Tensor],
                best_layers: Mapping[Hashable, str], method: str = "spearman", bootstrap: bool = False,
                n_bootstrap: int = 1000,
                selection_scores: Optional[Mapping[Hashable, Sequence[Dict]]] = None) -> List[Dict]:
    """The per-(region, subject) scoring block of `_eval_rsa` / `_eval_rsa_nsd_synthetic`
    (evals.py:333-398, :489-546) for all pairs in one device job.

    model_features: {layer: (n_test, ...)} exact activations of every distinct best layer (host or device);
    neural_test:    {pair key: (n_test, V)} test responses, iterated in the caller's (region, subject) order;
    best_layers:    {pair key: layer name}.
    Returns one dict per pair, in `neural_test` order, with the keys of evals.py:375-386:
    layer, compare_method, score, ci_low, ci_high, analysis, layer_selection_scores[, bootstrap_scores].
    """
    method = method.lower()
    keys = list(neural_test.keys())
    layers = sorted({best_layers[k] for k in keys})            # the reference re-extracts sorted(unique layers), :309
    col = {name: i for i, name in enumerate(layers)}
    feats = [rsa._flat(model_features[name]) for name in layers]
    neur = [torch.as_tensor(neural_test[k], dtype=torch.float32) for k in keys]
    res = rsa.rsa_grouped(feats, neur, method, bootstrap=bootstrap, n_bootstrap=n_bootstrap, seed=SELECTION_SEED)
    out = []
    for j, k in enumerate(keys):
        i = col[best_layers[k]]
        result = {"layer": best_layers[k], "compare_method": method, "score": float(res["point"][i, j]),
                  "ci_low": None, "ci_high": None, "analysis": "rsa",
                  "layer_selection_scores": list(selection_scores[k]) if selection_scores is not None else []}
        if bootstrap:
            result["ci_low"] = float(res["ci_low"][i, j])
            result["ci_high"] = float(res["ci_high"][i, j])
            result["bootstrap_scores"] = res["scores"][i, j].tolist()
        out.append(result)
    return out


def eval_rsa_scores(train_acts_by_pair: Mapping[Hashable, Mapping[str, torch.Tensor]],
                    train_neural_by_pair: Mapping[Hashable, torch.Tensor],
                    test_features: Mapping[str, torch.Tensor], test_neural_by_pair: Mapping[Hashable, torch.Tensor],
                    method: str = "spearman", n_select: Optional[int] = 1000, bootstrap: bool = False,
                    n_bootstrap: int = 1000) -> List[Dict]:
    """Phase 1 + scoring of `_eval_rsa` (evals.py:249-287 + :333-398) from already extracted activations:
    per-pair layer selection on the training split, then every pair's best layer scored on the test split.
    `test_features` must hold the exact test activations of every layer that can be selected."""
    best, sel_scores = {}, {}
    for k in test_neural_by_pair.keys():
        best[k], sel_scores[k] = select_layer(train_acts_by_pair[k], train_neural_by_pair[k], method, n_select)
    return score_pairs(test_features, test_neural_by_pair, best, method, bootstrap, n_bootstrap, sel_scores)


def _compare_rdms(rdm1: torch.Tensor, rdm2: torch.Tensor, correlation: str) -> float:
    """`experiments/binary_pc_rsa/main.py:28-38`: the same comparison with a positional method name."""
    return rsa.compute_rdm_correlation(rdm1, rdm2, correlation=correlation)


# --------------------------------------------------------------------------------------------------
# Signature-level drop-ins for `_eval_rsa` (visreps/evals.py:222-398) and `_eval_rsa_nsd_synthetic` (:442-548)
# --------------------------------------------------------------------------------------------------
# The reference functions reach everything around the numerics through names of their own module
# (`mutils`, `get_transform`, `_make_loader`, `reconstruct_from_pcs`, `save_results`, `OmegaConf`, `rprint`, `pd`,
# `load_nsd_synthetic_test_data`, `_lookup_nsd_best_layers`).  The replacements below resolve the same names on the
# *host* module at call time, so the producer, the loaders and the results store stay the reference's own (or
# whatever the user patched them with); only the loops around the RSA numerics change.
import sys


class _Host:
    """Name resolution on the module the drop-in is bound into (normally `visreps.evals`)."""

    def __init__(self, module):
        self._m = module

    def __getattr__(self, name):
        try:
            return getattr(self._m, name)
        except AttributeError:
            raise AttributeError(f"{getattr(self._m, '__name__', self._m)!r} has no attribute {name!r}, which the "
                                 f"visreps_b200 drop-in for _eval_rsa needs (see INTEGRATION.md)") from None

    def say(self, msg, style="info"):
        fn = getattr(self._m, "rprint", None)
        if fn is not None:
            fn(msg, style=style)

    def merged_cfg(self, cfg, extra):
        oc = getattr(self._m, "OmegaConf", None)
        if oc is not None:
            return oc.merge(cfg, extra)
        out = dict(cfg)
        out.update(extra)
        return out

    def frame(self, rows):
        pd = getattr(self._m, "pd", None)
        if pd is None:
            import pandas as pd
        return pd.DataFrame(rows)


def _host_module(host=None):
    if host is not None:
        return host if isinstance(host, _Host) else _Host(host)
    mod = sys.modules.get("visreps.evals")
    if mod is None:
        raise RuntimeError("visreps.evals is not imported: call visreps_b200.install() after importing the reference, "
                           "or pass host=<module providing mutils, get_transform, _make_loader, save_results, ...>")
    return _Host(mod)


def _test_neural_tensor(test_neural_dict, test_ids, filter_missing=True):
    """evals.py:341-343 / :493-495: stack the pair's test responses in shared-test-id order."""
    if filter_missing:
        responses = [test_neural_dict[sid] for sid in test_ids if sid in test_neural_dict]
    else:
        responses = [test_neural_dict[sid] for sid in test_ids]
    return torch.as_tensor(np.stack(responses).squeeze(), dtype=torch.float32)


def _pair_result(best_layer, method, point, bootstrap, ci_low, ci_high, scores, selection_scores):
    """The result dictionary of evals.py:375-386 / :531-542 (same keys, same order, same value types)."""
    result = {"layer": best_layer, "compare_method": method, "score": float(point),
              "ci_low": float(ci_low) if bootstrap else None, "ci_high": float(ci_high) if bootstrap else None,
              "analysis": "rsa", "layer_selection_scores": selection_scores}
    if bootstrap:
        result["bootstrap_scores"] = scores.tolist()
    return result


def _score_and_save(host, cfg, regions, subjects, best_layers, selection_scores, model_rdms, neural_tensors, method,
                    bootstrap, n_bootstrap):
    """Scoring + save block shared by both drop-ins (evals.py:333-398, :489-546): every (region, subject) pair against
    its best layer's RDM, in ONE grouped device job (the reference re-seeds RandomState(42) per pair, :356, so all
    pairs see the same index sets and the rankings are shared)."""
    keys = [(region, subj) for region in regions for subj in subjects]
    layers = sorted(model_rdms.keys())
    col = {name: i for i, name in enumerate(layers)}
    dev = rsa._lib.require_cuda()
    flags = torch.zeros(len(keys), dtype=torch.int32, device=dev)
    neural_rdms = [rsa._rdm_device(neural_tensors[k], dev, 1e-12, flags, j) for j, k in enumerate(keys)]
    sizes = {int(r.shape[0]) for r in neural_rdms} | {int(model_rdms[l].shape[0]) for l in layers}
    if len(sizes) != 1:
        raise ValueError("RDMs must share the same 2-D shape")
    if bootstrap:
        res = rsa.bootstrap_rdm_correlation_grouped([model_rdms[l] for l in layers], neural_rdms, method,
                                                    n_bootstrap=n_bootstrap, seed=SELECTION_SEED)
        ci_low, ci_high = rsa.percentile_ci(res["scores"])
    else:
        n = sizes.pop()
        if n <= 1:
            res = {"point": np.full((len(layers), len(keys)), np.nan)}
        else:
            masks = rsa.build_masks(None, n, dev, with_full=True)
            prepared = rsa._prepare_many([model_rdms[l] for l in layers] + neural_rdms, method)
            out = rsa.compare_grouped(prepared[:len(layers)], prepared[len(layers):], method, masks)
            res = {"point": out[:, :, 0].cpu().numpy()}
    all_results, to_save = [], []
    current_region = None
    for j, (region, subj) in enumerate(keys):
        if region != current_region:
            host.say(f"\n  -- Region: {region} --", style="info")
            current_region = region
        best_layer = best_layers[region][subj]
        i = col[best_layer]
        result = _pair_result(best_layer, method, res["point"][i, j], bootstrap,
                              ci_low[i, j] if bootstrap else None, ci_high[i, j] if bootstrap else None,
                              res["scores"][i, j] if bootstrap else None, selection_scores[region][subj])
        msg = f"    subj {subj} | {method.capitalize():<10}| {best_layer} = {result['score']:.4f}"
        if bootstrap:
            msg += f"  [95% CI: {result['ci_low']:.4f}, {result['ci_high']:.4f}]"
        host.say(msg, style="highlight")
        if cfg.get("log_expdata"):
            to_save.append((host.frame([result]), host.merged_cfg(cfg, {"subject_idx": subj, "region": region})))
        all_results.append(result)
    if to_save:
        sink = RESULTS_SINK or (lambda pairs: [host.save_results(df, c) for df, c in pairs])
        sink(to_save)
    return host.frame(all_results)


RESULTS_SINK = None       # set by install(batched_sink=True): visreps_b200.results_sink.save_results_batch


def _eval_rsa(cfg, model, acts, ids, all_data, subjects, regions, dev, verbose, host=None):
    """Drop-in for `_eval_rsa` (visreps/evals.py:222-398), same arguments and the same DataFrame.

    Phase 1 (evals.py:249-287): per (region, subject) layer selection on the SRP activations.  Pairs whose aligned
    training rows and selection subsample coincide (the regions of one subject) share the model-side RDMs: one grouped
    job = all layers x those regions' neural RDMs, instead of one RDM build per layer per pair.
    Phase 2 (evals.py:292-326): exact re-extraction of the distinct best layers through the host's
    `mutils.extract_single_layer` (kept on the device if the producer hands device tensors), one RDM per layer.
    Scoring (evals.py:333-398): all pairs x their best layers x (1 + n_bootstrap) masks in one grouped device job.
    `acts` may live on the host (reference producer, models/utils.py:344) or on the device.
    """
    host = _host_module(host)
    from .alignment import _align_stimulus_level
    method = cfg.get("compare_method", "spearman").lower()
    bootstrap = cfg.get("bootstrap", False)
    n_bootstrap = cfg.get("n_bootstrap", 1000)
    n_select = cfg.get("n_select", 1000)
    neural = all_data["neural"]
    shared_test_ids = all_data["shared_test_ids"]
    stimuli = all_data["stimuli"]

    # ---- phase 1 -------------------------------------------------------------------------------
    host.say("\n  Phase 1: Per-subject layer selection", style="info")
    per_region_layers = {region: {} for region in regions}
    per_region_scores = {region: {} for region in regions}
    names = list(acts.keys())
    groups: Dict[tuple, dict] = {}
    for region in regions:
        for subj in subjects:
            targets = neural[region][subj]["train"]
            idx = [i for i, k in enumerate(ids) if str(k) in targets]           # alignment.py:29
            if idx:
                train_neural = torch.as_tensor(np.stack([targets[str(ids[i])] for i in idx]), dtype=torch.float32)
            else:
                train_neural = torch.empty(0, dtype=torch.float32)
            n_train = int(train_neural.size(0)) if train_neural.ndim else 0
            sel_idx = selection_indices(n_train, n_select)
            rows = np.asarray(idx, dtype=np.int64)[sel_idx] if n_train else np.zeros(0, dtype=np.int64)
            key = rows.tobytes()
            g = groups.setdefault(key, {"rows": rows, "pairs": [], "neural": []})
            g["pairs"].append((region, subj))
            g["neural"].append(train_neural[torch.from_numpy(np.asarray(sel_idx, dtype=np.int64))] if n_train else train_neural)
    for g in groups.values():
        rows = torch.from_numpy(g["rows"])
        if len(rows) == 0:
            point = np.full((len(names), len(g["pairs"])), np.nan)
        else:
            feats = [rsa._flat(acts[k][rows.to(acts[k].device)]) for k in names]
            point = rsa.rsa_grouped(feats, g["neural"], method, bootstrap=False)["point"]
        for j, (region, subj) in enumerate(g["pairs"]):
            scores, best_layer, best_score = [], None, -float("inf")
            for name, s in zip(names, point[:, j]):
                s = float(s)
                scores.append({"layer": name, "score": s})
                if s > best_score:
                    best_score, best_layer = s, name
            per_region_layers[region][subj] = best_layer
            per_region_scores[region][subj] = scores
            if verbose:
                host.say(f"    {region} subj {subj}: {best_layer} ({best_score:.4f}), {len(g['rows'])} stimuli for selection")
    del groups
    host.say("  Freed bulk SRP activations", style="success")

    # ---- phase 2: exact test RDMs of the distinct best layers ----------------------------------------
    host.say("\n  Phase 2: Test evaluation", style="info")
    test_stimuli = {sid: stimuli[sid] for sid in shared_test_ids if sid in stimuli}
    dl_test = host._make_loader(test_stimuli, host.get_transform(ds_stats="imgnet"), cfg.batchsize, cfg.num_workers)
    host.say(f"  Test dataloader: {len(test_stimuli)} stimuli", style="success")
    unique_layers = set()
    for region_layers in per_region_layers.values():
        unique_layers.update(region_layers.values())
    model_rdms = _exact_model_rdms(host, cfg, model, dl_test, dev, unique_layers, shared_test_ids, "Re-extracting")
    del dl_test

    # ---- scoring ---------------------------------------------------------------------------------
    neural_tensors = {(region, subj): _test_neural_tensor(neural[region][subj]["test"], shared_test_ids)
                      for region in regions for subj in subjects}
    return _score_and_save(host, cfg, regions, subjects, per_region_layers, per_region_scores, model_rdms, neural_tensors,
                           method, bootstrap, n_bootstrap)


def _exact_model_rdms(host, cfg, model, dl_test, dev, unique_layers, test_ids, verb):
    """evals.py:312-322 / :473-482: one exact RDM per distinct best layer, kept on the CUDA device."""
    pca_k = cfg.get("pca_k", 1)
    cuda = rsa._lib.require_cuda()
    flags = torch.zeros(max(1, len(unique_layers)), dtype=torch.int32, device=cuda)
    model_rdms = {}
    for slot, layer in enumerate(sorted(unique_layers)):
        host.say(f"  {verb} {layer} without SRP...", style="info")
        exact_acts, _ = host.mutils.extract_single_layer(model, dl_test, dev, layer, test_ids)
        if cfg.get("reconstruct_from_pcs"):
            exact_acts = host.reconstruct_from_pcs({layer: exact_acts}, pca_k)[layer]
            host.say(f"    Reconstructed from {pca_k} PCs", style="info")
        model_rdms[layer] = rsa._rdm_device(exact_acts, cuda, 1e-12, flags, slot)
        del exact_acts
    return model_rdms


def _eval_rsa_nsd_synthetic(cfg, subjects, regions, dev, verbose, host=None):
    """Drop-in for `_eval_rsa_nsd_synthetic` (visreps/evals.py:442-548): best layers looked up by the host, the distinct
    layers' exact RDMs on the 220 synthetic stimuli, then all (region, subject) pairs scored in one grouped device job."""
    host = _host_module(host)
    method = cfg.get("compare_method", "spearman").lower()
    bootstrap = cfg.get("bootstrap", False)
    n_bootstrap = cfg.get("n_bootstrap", 1000)
    best_layers = host._lookup_nsd_best_layers(cfg, subjects, regions)
    if verbose:
        for region in regions:
            for subj in subjects:
                host.say(f"    {region} subj {subj}: reusing layer {best_layers[region][subj]} from NSD", style="info")
    test_data = host.load_nsd_synthetic_test_data(cfg, subjects=subjects, regions=regions)
    test_ids, stimuli, neural = test_data["test_ids"], test_data["stimuli"], test_data["neural"]
    host.say(f"  Loaded {len(test_ids)} synthetic test stimuli", style="success")
    model = host.mutils.load_model(cfg, dev, verbose=verbose)
    model = host.mutils.configure_feature_extractor(cfg, model, verbose=verbose)
    dl_test = host._make_loader(stimuli, host.get_transform(ds_stats="imgnet"), cfg.batchsize, cfg.num_workers)
    unique_layers = {l for rl in best_layers.values() for l in rl.values()}
    model_rdms = _exact_model_rdms(host, cfg, model, dl_test, dev, unique_layers, test_ids, "Extracting")
    del model, dl_test
    neural_tensors = {(region, subj): _test_neural_tensor(neural[region][subj], test_ids, filter_missing=False)
                      for region in regions for subj in subjects}
    no_scores = {region: {subj: [] for subj in subjects} for region in regions}
    return _score_and_save(host, cfg, regions, subjects, best_layers, no_scores, model_rdms, neural_tensors, method,
                           bootstrap, n_bootstrap)
