"""Kendall tau-a bootstrap timing: one (layer, ROI) pair, B replicates."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import visreps_b200 as vb
from visreps_b200 import synth, _lib
dev = torch.device("cuda", 0)
sizes = [int(x) for x in os.environ.get("KN", "220,1000").split(",")]
for n, B in [(n, 200 if n < 500 else int(os.environ.get("QB", 100))) for n in sizes]:
    z = synth.latent(n, dev, 1)
    a = vb.compute_rdm(synth.layer_features(z, 4096, 11))
    b = vb.compute_rdm(z @ torch.randn(synth.LATENT, 500, device=dev) + 0.5 * torch.randn(n, 500, device=dev))
    for rep in range(2):
        l0 = _lib.launch_count()
        torch.cuda.synchronize(); t = time.perf_counter()
        p, lo, hi, s = vb.bootstrap_rdm_correlation(a, b, "kendall", n_bootstrap=B)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"kendall n={n} B={B}: {dt*1e3:.1f} ms -> {B/dt:.0f} replicates/s, {_lib.launch_count()-l0} launches, tau={p:.4f} [{lo:.4f}, {hi:.4f}]", flush=True)
