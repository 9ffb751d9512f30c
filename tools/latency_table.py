"""Latency of the drop-in entry points at the NSD shape (n = 1000), warm, device-resident inputs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import visreps_b200 as vb
from visreps_b200 import synth, evals as vevals
dev = torch.device("cuda", 0)
n = 1000
feats, neur = synth.nsd_workload(n, dev, seed=0)
X = [x for _, x in feats]; Y = [y for _, y in neur]
def tm(label, fn, reps=5):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    print(f"{label:72s} {1e3 * (time.perf_counter() - t) / reps:9.3f} ms", flush=True)
    return r
fc7 = X[6]
tm("compute_rdm(fc7 4096-d)", lambda: vb.compute_rdm(fc7))
tm("compute_rdm(conv1 193,600-d)", lambda: vb.compute_rdm(X[0]))
a, b = vb.compute_rdm(fc7), vb.compute_rdm(Y[0])
for m in ("Spearman", "Pearson", "Kendall"):
    tm(f"compute_rdm_correlation({m})", lambda m=m: vb.compute_rdm_correlation(a, b, correlation=m))
for m in ("spearman", "pearson", "kendall"):
    tm(f"bootstrap_rdm_correlation({m}, 1000 replicates)", lambda m=m: vb.bootstrap_rdm_correlation(a, b, m, n_bootstrap=1000), reps=2)
tm("rsa_grouped(8 layers x 8 ROIs, spearman, no bootstrap)  [BASELINE config 2]", lambda: vb.rsa_grouped(X, Y, "spearman", bootstrap=False))
tm("rsa_grouped(8 layers x 8 ROIs, spearman, 1000 replicates)", lambda: vb.rsa_grouped(X, Y, "spearman", n_bootstrap=1000))
tm("rsa_grouped(8 layers x 8 ROIs, pearson, 1000 replicates)", lambda: vb.rsa_grouped(X, Y, "pearson", n_bootstrap=1000))
acts = {f"l{i}": x for i, x in enumerate(X)}
tm("evals.select_layer(8 layers, one ROI, n_select=1000)", lambda: vevals.select_layer(acts, Y[0], "spearman", 1000))
tr = vb.AlignmentData({k: v[:800] for k, v in acts.items()}, Y[0][:800], [str(i) for i in range(800)])
te = vb.AlignmentData({k: v[800:] for k, v in acts.items()}, Y[0][800:], [str(i) for i in range(800, 1000)])
tm("compute_rsa(800 train / 200 test, 8 layers, 1000 replicates)", lambda: vb.compute_rsa({"compare_method": "spearman"}, tr, te, n_bootstrap=1000), reps=3)
from visreps_b200 import encoding as venc
yt = Y[7]                                                   # (1000, 7600) "ventral" responses as targets
yp = 0.5 * yt + torch.randn_like(yt)
tm("encoding.bootstrap_encoding_score(1000 x 7600 predictions, 1000 replicates)", lambda: venc.bootstrap_encoding_score(yt, yp, n_bootstrap=1000), reps=3)
