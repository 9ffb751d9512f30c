"""Small end-to-end case for compute-sanitizer (all kernels, tiny sizes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import visreps_b200 as vb
dev = torch.device("cuda", 0)
rs = np.random.RandomState(0)
n = 70
x = torch.from_numpy(np.maximum(rs.standard_normal((n, 150)), 0).astype(np.float32)).to(dev)
y = torch.from_numpy(rs.standard_normal((n, 45)).astype(np.float32)).to(dev)
a, b = vb.compute_rdm(x), vb.compute_rdm(y)
vb.compute_rdm(x, correlation="Spearman")
for m in ("spearman", "pearson", "kendall"):
    print(m, vb.bootstrap_rdm_correlation(a, b, m, n_bootstrap=3)[0])
res = vb.rsa_grouped([x.cpu(), x[:, :40].cpu()], [y.cpu(), y[:, :20].cpu(), y[:, 5:30].cpu()], "spearman", n_bootstrap=5)
print(res["point"])
torch.cuda.synchronize()
print("sanitize case ok")
