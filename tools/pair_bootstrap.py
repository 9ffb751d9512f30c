"""One (model RDM, neural RDM) pair, B replicates: wall / device time and stage breakdown of bootstrap_rdm_correlation
(the call of the secondary callers: run_reconstruction.py:137-158, binary_pc_rsa/main.py:28-38)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import visreps_b200 as vb
from visreps_b200 import synth, _lib
dev = torch.device("cuda", 0)
n = int(os.environ.get("QN", 1000)); B = int(os.environ.get("QB", 1000))
z = synth.latent(n, dev, 1)
a = vb.compute_rdm(synth.layer_features(z, 4096, 11))
b = vb.compute_rdm(synth.roi_responses(z, 2000, 500))
for method in ("spearman", "pearson", "kendall"):
    for rep in range(3):
        if rep == 2: _lib.profile_start()
        torch.cuda.synchronize(); t = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        l0 = _lib.launch_count()
        res = vb.bootstrap_rdm_correlation(a, b, method, n_bootstrap=B)
        e1.record(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    prof = _lib.profile_stop()
    print(f"{method:9s} n={n} B={B}: wall {dt*1e3:.2f} ms, device {e0.elapsed_time(e1):.2f} ms, {_lib.launch_count()-l0} launches; stages:",
          {k: round(v["ms"], 2) for k, v in prof.items() if v["ms"] > 0})
