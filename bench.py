#!/usr/bin/env python
"""bench.py -- bootstrap RSA replicates/s (and RDM TFLOP/s) on B200, next to the reference CPU path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (oracle port) on host cores

Workload (BASELINE.json metric / north_star target): NSD shared-1000 shape --
n = 1000 stimuli, one model = 8 AlexNet layers (conv1..fc8, 494,184 features in total) x 8 ROI
response matrices, Pearson RDMs + Spearman comparison, 1000 bootstrap replicates per (layer, ROI) pair.
One step = one whole pass of the hot path over that input: 16 RDMs, 16 presorts, 64 point estimates,
64 x 1000 replicate comparisons, percentile CIs.  value = replicate comparisons per second.
With N GPUs every rank scores its own model (same ROIs) -- the granularity-sweep sharding of
BASELINE config 4 -- and the per-replicate score vectors are combined with one NCCL all-gather per
step (weak scaling; value = all ranks' replicates / max-over-ranks time).
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "bootstrap_rsa_replicates_per_s"
UNIT = "replicates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1000, help="stimuli")
    ap.add_argument("--n-bootstrap", type=int, default=1000)
    ap.add_argument("--small", action="store_true", help="shrunken layer widths (smoke runs only; not a valid bench)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling THINGS-shape job (BASELINE config 3)")
    ap.add_argument("--strong-steps", type=int, default=5)
    return ap.parse_args()


def workload_dims(args):
    from visreps_b200 import synth
    layers, rois = dict(synth.ALEXNET_DIMS), dict(synth.NSD_ROIS)
    if args.small:
        layers = {k: max(64, v // 64) for k, v in layers.items()}
        rois = {k: max(32, v // 8) for k, v in rois.items()}
    return layers, rois


def config_dict(args, layers, rois, world):
    return {
        "workload": f"NSD shared-1000 shape: n={args.n}, {len(layers)} AlexNet layers x {len(rois)} ROIs, "
                    f"Pearson RDM + Spearman compare, {args.n_bootstrap} bootstrap replicates per pair"
                    + (" [SMALL smoke variant]" if args.small else ""),
        "n_stimuli": args.n, "layers": layers, "rois": rois, "n_bootstrap": args.n_bootstrap,
        "pairs_per_gpu": len(layers) * len(rois), "sharding": f"one model per GPU x{world}, 1 all-gather of scores per step",
        "l2_policy": "inputs (2.05 GB of features per model) exceed the 126 MB L2; no flush needed",
        "index_draws": "replayed from RandomState(seed) inside every timed step (rsa.MASK_CACHE off: nothing reused across steps)",
    }


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline (host cores; the only place bench.py executes oracle/)
# ----------------------------------------------------------------------------------------------
def reference_sampler(args, layers, rois):
    from oracle import cpu_bench
    return cpu_bench.ReferenceSampler(args.n, layers, rois, args.n_bootstrap)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    layers, rois = workload_dims(args)
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    smp = reference_sampler(args, layers, rois)
    t_rdm = smp.time_rdms()                      # measured once (seconds-long); reused by every step
    for _ in range(args.warmup):
        smp.step(t_rdm)
    vals, jobs, per = [], [], []
    for _ in range(args.steps):
        r = smp.step(t_rdm)
        vals.append(r["value"])
        jobs.append(r["job_seconds"])
        per.append(r["per_compare_s_per_core"])
    as_written = smp.as_written(t_rdm)
    smp.close()
    per_core = float(np.mean(per))
    v = float(np.mean(vals))
    line = {
        "metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(jobs)) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 RDM (CPU sgemm) / f64 Spearman (SciPy)", "data": "synthetic",
        "config": config_dict(args, layers, rois, world),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": smp.cores, "kind": "port", "sample": smp.sample_description(),
                         "t_rdm_s": t_rdm, "s_per_replicate_per_core": per_core, "replicates_per_s_per_core": 1.0 / per_core,
                         "as_written": as_written},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "ms_per_step is the whole-job time scaled from the bounded sample; a CPU job does not speed up with --gpus",
    }
    emit(line)


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(gpu_index), "-lms", "100"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        for ln in self.tmp.read().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.tmp.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    layers, rois = workload_dims(args)

    # CPU baseline first: it forks worker processes, which must happen before CUDA is initialised
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        smp = reference_sampler(args, layers, rois)
        r = smp.step()
        cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": smp.cores, "kind": "port",
                        "sample": smp.sample_description(), "t_rdm_s": r["t_rdm"],
                        "s_per_replicate_per_core": r["per_compare_s_per_core"],
                        "replicates_per_s_per_core": 1.0 / r["per_compare_s_per_core"],
                        "as_written": smp.as_written(r["t_rdm"])}
        smp.close()
        torch.set_num_threads(max(1, min(8, smp.cores)))

    import torch.distributed as dist
    import visreps_b200 as vb
    from visreps_b200 import _lib, synth
    from visreps_b200 import rsa as _vrsa
    _vrsa.MASK_CACHE = False          # every step replays the reference's index draws; the product keeps the last few (seed-keyed)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    tf32_peak_info = measure_tf32_peak(dev) if (rank == 0 and not args.small) else None
    feats, neur = synth.nsd_workload(args.n, dev, seed=rank, layers=layers, rois=rois)
    X = [x for _, x in feats]
    Y = [y for _, y in neur]
    nA, nB, B = len(X), len(Y), args.n_bootstrap
    in_bytes = sum(t.numel() * 4 for t in X + Y)
    neural_bytes = sum(t.numel() * 4 for t in Y)
    flops_step = sum(args.n * (args.n + 1) * t.shape[1] for t in X + Y)
    gathered = torch.empty((world, nA, nB, B + 1), dtype=torch.float64, device=dev) if world > 1 else None

    def step_device():
        res = vb.rsa_grouped(X, Y, "spearman", bootstrap=True, n_bootstrap=B, seed=42)
        if world > 1:   # the one exchange step of the path: per-replicate score vectors, replicate order kept
            mine = torch.from_numpy(np.concatenate([res["point"][:, :, None], res["scores"]], axis=2)).to(dev)
            dist.all_gather_into_tensor(gathered, mine)
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        barrier()
        if profile:
            _lib.profile_start()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        last = None
        for _ in range(steps):
            last = fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        prof = _lib.profile_stop() if profile else None
        launches = _lib.launch_count() - l0
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), prof, launches, last

    # the sampler is started before the warm-up so that nvidia-smi's own start-up (driver init, ~1 s of host
    # and driver-lock contention) is over before the timed region; it then only polls every 100 ms
    clocks = ClockSampler(local_rank) if (rank == 0 and not os.environ.get('VR_BENCH_NO_CLOCKS')) else None
    if clocks is not None:
        time.sleep(1.5)
    # the imports leave ~1.5 M long-lived Python objects; a full collection of that heap costs ~34 ms here, more
    # than a whole step.  Park them in the permanent generation so later collections only see the steps' garbage.
    gc.collect()
    gc.freeze()
    for _ in range(args.warmup):
        step_device()
    ms_dev, wall_dev, prof, launches, res = timed(step_device, args.steps, profile=True)
    clock_info = clocks.stop() if clocks else None
    # The timed steps build the 16 RDMs on two alternating streams (the row-centring pass of one matrix overlaps the SYRK of
    # the previous one), so their per-family event times overlap.  The per-kernel breakdown and the SYRK roofline come from a
    # second, untimed pass with the builds serialised; the dominant kernel (the sweep) is timed in the timed region itself.
    _vrsa.RDM_OVERLAP = False
    step_device()
    serial_steps = max(1, min(args.steps, 5))
    _ms_ser, _wall_ser, prof_serial, _, _ = timed(step_device, serial_steps, profile=True)
    _vrsa.RDM_OVERLAP = True

    # ---- end to end: host (pinned) inputs -> public API -> host results, copies inside the timed region
    Xh = [x.cpu().pin_memory() for x in X]
    Yh = [y.cpu().pin_memory() for y in Y]

    def step_e2e():
        return step_device_host(Xh, Yh)

    def step_device_host(xh, yh):
        r = vb.rsa_grouped(xh, yh, "spearman", bootstrap=True, n_bootstrap=B, seed=42)
        if world > 1:
            mine = torch.from_numpy(np.concatenate([r["point"][:, :, None], r["scores"]], axis=2)).to(dev)
            dist.all_gather_into_tensor(gathered, mine)
        return r

    step_e2e()
    ms_e2e, wall_e2e, _, _, res_e2e = timed(step_e2e, args.steps)
    assert np.array_equal(res["scores"], res_e2e["scores"]), "host-input and device-input runs disagree"
    assert np.isfinite(res["scores"]).all() and np.isfinite(res["point"]).all()

    # ---- end to end with a device-resident producer (visreps_b200.producer / install(device_producer=True)): the model
    # activations never leave the GPU (models/utils.py:344,391 bypassed); the neural responses still come from host
    # memory (the loaders of dataloaders/neural.py return numpy arrays) and the results still go back to the host.
    def step_e2e_device_producer():
        return step_device_host(X, Yh)

    step_e2e_device_producer()
    ms_dp, wall_dp, _, _, res_dp = timed(step_e2e_device_producer, args.steps)
    assert np.array_equal(res["scores"], res_dp["scores"])
    del Xh

    # ---- parity gates (SURVEY 8d / BASELINE.md 3.6), part of the cpu_baseline leg: the oracle is the checker only
    if rank == 0 and cpu_baseline is not None:
        cpu_baseline["parity_gates"] = parity_gates(vb, X, Y, res, B, dev)

    # ---- strong scaling (BASELINE config 3): THINGS shape, 1854 concepts, behavioural RDM (d = 66) vs 8 layers,
    # B replicates of ONE job sharded over the ranks (dist.rsa_grouped_sharded: sharded RDM builds + presorts, blob
    # all-gather over NVLink, replicate shards, one all-gather of the scores in replicate order)
    strong = None
    if not args.no_strong and not args.small:
        from visreps_b200 import dist as vdist
        del X, Y, feats, neur
        torch.cuda.empty_cache()
        n3 = 1854
        z3 = synth.latent(n3, dev, 1)                                # same seeds on every rank: identical inputs
        feats3 = [synth.layer_features(z3, d, 10 + i) for i, d in enumerate(layers.values())]
        g3 = torch.Generator(device=dev).manual_seed(7)
        beh = torch.relu(z3 @ torch.randn(synth.LATENT, 66, device=dev, generator=g3))

        def step_strong():
            if world > 1:
                return vdist.rsa_grouped_sharded(feats3, [beh], "spearman", n_bootstrap=B, seed=42)
            return vb.rsa_grouped(feats3, [beh], "spearman", bootstrap=True, n_bootstrap=B, seed=42)

        for _ in range(2):
            step_strong()
        ms_s, wall_s, _, _, res_s = timed(step_strong, args.strong_steps)
        t_s = max(ms_s, wall_s) / args.strong_steps
        checksum = float(np.sum(res_s["scores"])) + float(np.sum(res_s["point"]))
        strong = {"workload": f"THINGS shape (BASELINE config 3): n={n3}, behavioural RDM d=66 x {len(feats3)} layers, {B} replicates of ONE job",
                  "scaling": "strong", "n_gpus": world, "ms_per_step": t_s, "value": len(feats3) * B / (t_s * 1e-3), "unit": UNIT,
                  "steps": args.strong_steps, "score_checksum": checksum,
                  "note": "wide layers are K-split over the ranks (balanced RDM builds), so RDM entries differ from the one-GPU build in the "
                          "last bits and score_checksum agrees across N to ~1e-6 of its value; rsa_grouped_sharded(ksplit=False) is bit-identical"}
        del feats3, beh

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # a step's host work (index draws, percentiles) overlaps device work; the step time is the max
    # of the device-event time and the wall time between the two synchronised barriers
    t_step_ms = max(ms_dev, wall_dev) / args.steps
    t_e2e_ms = max(ms_e2e, wall_e2e) / args.steps
    units = world * nA * nB * B
    value = units / (t_step_ms * 1e-3)
    e2e_value = units / (t_e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel family (device time from CUDA events inside the timed region)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    fam_ms = {k: v["ms"] for k, v in prof.items() if v["ms"] > 0}
    dom = max(fam_ms, key=fam_ms.get)
    n = args.n
    N = n * (n - 1) // 2
    m = int(n * 0.9)
    N_sub = m * (m - 1) // 2
    roofline = None
    if dom in ("rank", "rank_fused", "kept"):
        from visreps_b200 import rsa as vrsa_
        sweep = vrsa_.SPEARMAN_ENGINE == "sweep"
        G = (N + 31) // 32
        batches = (B + 1 + 31) // 32
        per_launch_ms = prof[dom]["ms"] / max(1, prof[dom]["scopes"])
        scopes_per_step = max(1.0, prof[dom]["scopes"] / args.steps)
        if sweep and dom == "rank_fused":
            kernel = f"sweep_kernel<{min(nA, 8)}, {min(nB, 8)}, compact>"
            # what ONE launch must read at least once: the count-table rows of its vectors ({kept word, count} per 32 sorted
            # positions, replicate and vector: 8 B each), their run words and the transposed masks; it writes (nA nB splits 32) sums
            compulsory = (nA + nB) * batches * (G + 1) * 256.0 + (nA + nB) * 4.0 * N + batches * n * 4.0
            rows_fetched = (nA + nB) * batches * float(N) * 256.0           # one 256-byte row per (pair, vector, batch): L2 -> SM
            what = ("compulsory bytes = the count-table rows + run words + masks the launch reads at least once from HBM; "
                    "the rows are then fetched from L2 once per (canonical pair, vector, 32-replicate batch)")
        else:
            fused = prof["rank_fused"]["scopes"] > 0
            vectors = {"rank": min(nA, nB, 8) if fused else nA + nB, "rank_fused": max(nA, nB), "kept": nA + nB}[dom]
            kernel = {"rank": "rank_kernel<materialise>", "rank_fused": "rank_kernel<fused dot>", "kept": "count kernels"}[dom]
            stride = (N + 31) // 32 * 32
            if dom == "rank":
                compulsory = vectors * (B + 1) * stride * 4.0          # the materialised dev vectors it writes
            elif dom == "rank_fused":
                compulsory = min(nA, nB, 8) * (B + 1) * stride * 4.0    # the materialised side, read once through L2
            else:
                compulsory = vectors * batches * (G + 1) * 256.0
            rows_fetched = None
            what = "compulsory bytes = the materialised centred-rank vectors (or count tables) the launch writes / reads once"
        achieved = compulsory / (per_launch_ms * 1e-3) / 1e9
        comparisons_per_launch = nA * nB * (B + 1) / scopes_per_step if dom == "rank_fused" else 0
        roofline = {"kernel": kernel, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": None, "peak_source": peak_src, "compulsory_bytes_per_launch": compulsory, "ms_per_launch": per_launch_ms,
                    "launches_per_step": scopes_per_step, "definition": what + "; achieved = compulsory bytes / CUDA-event time of the launch",
                    "work_sharing_equivalent_GBps": comparisons_per_launch * 8.0 * N_sub / (per_launch_ms * 1e-3) / 1e9 if comparisons_per_launch else None,
                    "work_sharing_equivalent_definition": "SURVEY 8d's 8*N_sub bytes per replicate comparison x the comparisons one launch finishes; NOT a "
                                                          "roofline fraction: nA*nB*(B+1) comparisons come from (nA+nB)*(B+1) rankings (evals.py:356)"}
        if rows_fetched:
            roofline["l2_row_GBps"] = rows_fetched / (per_launch_ms * 1e-3) / 1e9
            roofline["note"] = ("the sweep is not HBM bound: its table rows are L2 resident (one 64 MB batch block in flight, ncu lts hit rate 83 %, "
                                "5.7 GB of DRAM reads per launch); the binding resources are instruction issue with two warps per scheduler "
                                "(64 fp64 accumulators per lane = 255 registers) and the ~900-cycle L2 row latency "
                                "(profiles/r2_sweep_kernels_nsd_full.txt)")
        try:   # DRAM bytes per launch of this kernel from the committed ncu capture, if it is for this workload
            tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            wl = tr.get("workload", {})
            key = "sweep_kernel" if sweep and dom == "rank_fused" else roofline["kernel"]
            if (wl.get("n"), wl.get("nA"), wl.get("nB"), wl.get("masks")) == (n, nA, nB, B + 1) and key in tr:
                roofline["traffic"] = tr[key]["dram_bytes_per_launch"]
                roofline["traffic_source"] = tr[key].get("source", "ncu --set full")
        except Exception:
            pass
        if roofline["traffic"]:
            roofline["dram_GBps"] = roofline["traffic"] / (per_launch_ms * 1e-3) / 1e9
            roofline["dram_frac"] = roofline["dram_GBps"] / hbm_peak
    elif dom in ("rdm_syrk", "rdm_syrk_pair"):
        flops = flops_step
        tf_peak = tf32_peak_info["burst_tflops"] if tf32_peak_info else float(peaks.get("bf16_tflops", 1599.5)) / 2.0
        ach = 3.0 * flops * args.steps / ((prof["rdm_syrk"]["ms"] + prof["rdm_syrk_pair"]["ms"]) * 1e-3) / 1e12
        roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                    "traffic": None, "peak_source": "TF32 dense peak measured in this run" if tf32_peak_info else "bf16 peak / 2"}
    # second roofline object: the tensor-core SYRK family (all RDM builds of the step), useful TF32 work
    roofline_syrk = None
    syrk_ms = prof_serial.get("rdm_syrk", {}).get("ms", 0) + prof_serial.get("rdm_syrk_pair", {}).get("ms", 0)
    if syrk_ms > 0:
        tf32_useful = 3.0 * flops_step * serial_steps / (syrk_ms * 1e-3) / 1e12
        if tf32_peak_info:
            tf32_peak, tf32_src = tf32_peak_info["burst_tflops"], "measured in this run: torch.matmul fp32 with TF32 tensor cores (cuBLAS), 8192^3, best of 10"
        else:
            tf32_peak, tf32_src = float(peaks.get("bf16_tflops", 1599.5)) / 2.0, "half the measured dense bf16 rate (MEASURED_PEAKS.json bf16_tflops)"
        roofline_syrk = {"kernel": "syrk_tc2_kernel (pair tiles, cta_group::2) / syrk_tc_kernel", "bound": "tensor",
                         "achieved": tf32_useful, "peak": tf32_peak, "unit": "TFLOP/s", "frac": tf32_useful / tf32_peak,
                         "definition": "3 TF32 products x n(n+1)d useful flop per RDM (upper triangle) / CUDA-event time of the SYRK launches "
                                       "in the serialised pass (in the timed steps the row-centring pass of the next matrix shares the SMs with them)",
                         "peak_source": tf32_src,
                         "note": "ncu: sm__pipe_tensor_cycles_active 92.3 % for the pair-tile kernel at n=1000, d=193,600 "
                                 "(profiles/r1_syrk_tc2_conv1_full.txt); 11 % of the executed tensor work lies below the diagonal of the "
                                 "256-wide tiling and the short rows (ROIs, fc layers) are launch-bound"}
    rdm_flops = flops_step * serial_steps
    rdm_ms = syrk_ms + prof_serial["rdm_center"]["ms"] + prof_serial["rdm_epilogue"]["ms"]
    rdm_tflops = rdm_flops / (rdm_ms * 1e-3) / 1e12 if rdm_ms > 0 else None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 RDM / u32 ranks / i64+f64 Spearman", "data": "synthetic",
        "config": config_dict(args, layers, rois, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(in_bytes + B * ((n + 31) // 32) * 4),
                "d2h_bytes_per_step": int(nA * nB * (B + 1) * 8 + (nA + nB) * 4), "ms_per_step": t_e2e_ms},
        "e2e_device_producer": {"value": units / (max(ms_dp, wall_dp) / args.steps * 1e-3), "unit": UNIT,
                                "ms_per_step": max(ms_dp, wall_dp) / args.steps,
                                "h2d_bytes_per_step": int(neural_bytes + B * ((n + 31) // 32) * 4),
                                "d2h_bytes_per_step": int(nA * nB * (B + 1) * 8 + (nA + nB) * 4),
                                "definition": "model activations handed over on the device (visreps_b200.producer), neural responses "
                                              "from pinned host memory, results back to the host"},
        "strong_scaling": strong,
        "tf32_peak": tf32_peak_info,
        "gpu_launches": int(launches),
        "clocks": clock_info,
        "roofline": roofline,
        "roofline_syrk": roofline_syrk,
        "cpu_baseline": cpu_baseline,
        "rdm_tflops": rdm_tflops,
        "rdm_flops_definition": "n(n+1)d per RDM (upper triangle incl. diagonal, 2 flop/MAC), fp32-accurate",
        "stage_ms_per_step": {k: v["ms"] / serial_steps for k, v in prof_serial.items() if v["ms"] > 0},
        "stage_ms_definition": "CUDA-event time per kernel family in an untimed pass with the RDM builds on ONE stream (rsa.RDM_OVERLAP off), so "
                               "that family times do not overlap; the timed steps build the RDMs on two alternating streams and are "
                               f"{'%.2f' % ((_wall_ser if _wall_ser > _ms_ser else _ms_ser) / serial_steps - t_step_ms)} ms shorter than that pass",
        "stage_ms_per_step_timed_region": {k: v["ms"] / args.steps for k, v in prof.items() if v["ms"] > 0},
        "stage_launches_per_step": {k: v["launches"] / args.steps for k, v in prof.items() if v["launches"] > 0},
        "device_ms_per_step": ms_dev / args.steps, "wall_ms_per_step": wall_dev / args.steps,
        "score_sample": {"point_00": float(res["point"][0, 0]), "ci_00": [float(res["ci_low"][0, 0]), float(res["ci_high"][0, 0])]},
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def measure_tf32_peak(dev):
    """Dense TF32 tensor-core peak of this box ('peak(TF32) to be measured the same way' as the bf16 peak in
    MEASURED_PEAKS.json): cuBLAS through torch.matmul on fp32 operands with TF32 enabled, 8192^3, CUDA events; best of 10
    (burst) and back to back for ~1.5 s (sustained).  Library GEMM used as the yardstick only."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        m = 8192
        a = torch.randn(m, m, device=dev)
        b = torch.randn(m, m, device=dev)
        c = torch.empty(m, m, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(1500.0 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record(); e1.synchronize()
        flop = 2.0 * m ** 3
        return {"burst_tflops": flop / (best * 1e-3) / 1e12, "sustained_tflops": flop * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12,
                "how": "torch.matmul fp32 operands, allow_tf32=True (cuBLAS TF32 tensor cores), 8192^3, CUDA events"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def parity_gates(vb, X, Y, res, B, dev):
    """The gates SURVEY 8d asks to print with every benchmark, on this run's own inputs (CPU oracle = checker):
    RDM max-abs vs the port on one wide layer, ranks bit-exact, point + first two replicates vs the port."""
    from oracle import rsa_oracle as orc
    from visreps_b200 import rsa as vrsa
    li = min(3, len(X) - 1)                                   # conv4 (43,264 features) at the full shape
    n = int(X[li].shape[0])
    a_gpu, b_gpu = vb.compute_rdm(X[li]), vb.compute_rdm(Y[0])
    t0 = time.perf_counter()
    a_ref = orc.rdm(X[li].cpu())
    t_rdm = time.perf_counter() - t0
    rdm_err = float((a_gpu.cpu() - a_ref).abs().max())
    masks = vrsa.build_masks(None, n, dev, with_full=True)
    devv, _, _ = vrsa.subsample_rank(a_gpu, masks)
    N = n * (n - 1) // 2
    ranks_ok = bool(np.array_equal(devv[0].cpu().numpy().astype(np.int64) + (N + 1), orc.rank2x(orc.triu_vector(a_gpu))))
    idx = orc.draw_subsamples(n, 2, 42) if B >= 2 else None
    point_ref = orc.rdm_compare(a_gpu.cpu(), b_gpu.cpu(), correlation="Spearman")
    gates = {"layer": li, "rdm_max_abs_vs_port": rdm_err, "rdm_ok": rdm_err <= 1e-5, "ranks_bit_exact": ranks_ok,
             "point_abs_diff": abs(float(res["point"][li, 0]) - point_ref), "port_rdm_seconds": t_rdm}
    if idx is not None:
        want = orc.bootstrap_compare_with_indices(a_gpu.cpu(), b_gpu.cpu(), "spearman", idx)
        gates["replicates_max_abs_diff"] = float(np.abs(res["scores"][li, 0, :2] - np.asarray(want)).max())
    gates["scores_ok"] = max(gates["point_abs_diff"], gates.get("replicates_max_abs_diff", 0.0)) <= 1e-5
    assert gates["rdm_ok"] and gates["ranks_bit_exact"] and gates["scores_ok"], gates
    return gates


def _claim_stdout():
    """Everything that writes to fd 1 during the run (NCCL's version banner, library chatter) goes to stderr;
    the one JSON line is written to the real stdout at the end."""
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    return real


_REAL_STDOUT = None


def emit(line: dict) -> None:
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    global _REAL_STDOUT
    _REAL_STDOUT = _claim_stdout()
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
