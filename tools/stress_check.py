"""BASELINE config 5 (scale stress) feasibility: n = 10,000 stimuli (N = 49,995,000 pairs).
Checks properties that need no CPU oracle and prints stage timings.  SN/SD/SB env: n, d, replicates."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import visreps_b200 as vb
from visreps_b200 import rsa as vrsa, synth

dev = torch.device("cuda", 0)
n = int(os.environ.get("SN", 10000)); d = int(os.environ.get("SD", 193600)); B = int(os.environ.get("SB", 4))
def t(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"{label:42s} {1e3 * (time.perf_counter() - t0):10.1f} ms", flush=True); return r
z = synth.latent(n, dev, 1)
x = t(f"features n={n} d={d}", lambda: synth.layer_features(z, d, 11))
y = synth.roi_responses(z, 2000, 500)
a = t("compute_rdm(model) [cold: allocates scratch]", lambda: vb.compute_rdm(x))
a = t("compute_rdm(model)", lambda: vb.compute_rdm(x))
print(f"   -> {n*(n+1)*d/1e12:.2f} algorithmic TFLOP")
assert torch.equal(a, a.T) and float(a.diagonal().abs().max()) == 0.0
# spot-check 64 RDM rows (spread over the matrix) against the same formula in fp64 on the device (column chunks)
def fp64_rows(x, rows):
    nn, dd = x.shape
    s1 = torch.zeros(nn, dtype=torch.float64, device=dev); s2 = torch.zeros_like(s1)
    g = torch.zeros((len(rows), nn), dtype=torch.float64, device=dev)
    for c0 in range(0, dd, 16384):
        blk = x[:, c0:c0 + 16384].double()
        s1 += blk.sum(1); s2 += (blk * blk).sum(1); g += blk[rows] @ blk.T
    mean = s1 / dd
    cov = g / dd - mean[rows][:, None] * mean[None, :]
    sd = (s2 / dd - mean * mean).clamp_(min=0).add(1e-12).sqrt()
    sd[sd < 1e-11] = 1.0
    r = (cov / (sd[rows][:, None] * sd[None, :] + 1e-12)).clamp_(-1, 1)
    out = 1.0 - r
    out[torch.arange(len(rows), device=dev), rows] = 0.0
    return out
rows = torch.linspace(0, n - 1, 64, device=dev).long()
truth = t("fp64 truth of 64 RDM rows", lambda: fp64_rows(x, rows))
err = float((a[rows].double() - truth).abs().max())
print(f"   RDM max-abs error vs fp64 on 64 rows x {n} columns: {err:.3e}")
assert err <= 1e-5, err
del x, truth
b = t("compute_rdm(neural, V=2000)", lambda: vb.compute_rdm(y))
iu = torch.triu_indices(n, n, 1, device=dev)
for nm, m in (("model", a), ("neural", b)):
    v = m[iu[0], iu[1]]
    u, c = torch.unique(v, return_counts=True)
    print(f"   {nm} RDM: {v.numel()} entries in [{float(v.min()):.4f}, {float(v.max()):.4f}], {u.numel()} distinct values, "
          f"{100.0 * float(c[c > 32].sum()) / v.numel():.1f} % of the entries in runs longer than 32")
    del v, u, c
del iu
pa = t("presort(model)", lambda: vrsa.PresortedRDM(a))
pb = t("presort(neural)", lambda: vrsa.PresortedRDM(b))
idx = vrsa.draw_index_sets(n, B, 42)
masks = vrsa.build_masks(idx, n, dev, with_full=True)
from visreps_b200 import _lib
_lib.profile_start()
out = t(f"spearman: point + {B} replicates", lambda: vrsa.compare_grouped([pa], [pb], "spearman", masks))
print("   stages:", {k: round(v["ms"], 2) for k, v in _lib.profile_stop().items() if v["ms"] > 0})
self_ = t("spearman self-comparison", lambda: vrsa.compare_grouped([pa], [pa], "spearman", masks))
print("scores (first 6)", out[0, 0, :6].cpu().numpy())
assert np.allclose(self_.cpu().numpy(), 1.0, atol=1e-12)
# replicate 1 == full comparison of the gathered sub-RDMs
tix = torch.from_numpy(idx[0].astype(np.int64)).to(dev)
direct = vb.compute_rdm_correlation(a[tix][:, tix], b[tix][:, tix], correlation="Spearman")
print("replicate 0 vs direct sub-RDM comparison:", float(out[0, 0, 1]), direct)
assert abs(float(out[0, 0, 1]) - direct) < 1e-12
print("stress check ok; peak memory %.1f GB" % (torch.cuda.max_memory_allocated() / 2**30))
