"""Random small configurations of the public API against the oracle (development aid; the parity suite proper is tests/)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import visreps_b200 as vb
from oracle import rsa_oracle as orc
dev = torch.device("cuda", 0)
rs = np.random.RandomState(int(os.environ.get("SEED", 0)))
worst = 0.0
for trial in range(int(os.environ.get("TRIALS", 12))):
    n = int(rs.choice([int(v) for v in os.environ.get("SIZES", "5,9,33,64,65,97,130,257").split(",")]))
    nA, nB = int(rs.randint(1, 6)), int(rs.randint(1, 5))
    B = int(rs.choice([1, 3, 17, 33]))
    method = str(rs.choice(os.environ.get("METHODS", "spearman,pearson,kendall").split(",")))
    host = bool(rs.randint(2))
    quant = bool(rs.randint(3) == 0)
    lat = rs.standard_normal((n, 6)).astype(np.float32)
    def feat(d):
        x = lat @ rs.standard_normal((6, d)).astype(np.float32) + rs.standard_normal((n, d)).astype(np.float32)
        return np.round(x) if quant else x
    X = [torch.from_numpy(feat(int(rs.randint(3, 70)))) for _ in range(nA)]
    Y = [torch.from_numpy(feat(int(rs.randint(3, 40)))) for _ in range(nB)]
    Xi = X if host else [x.to(dev) for x in X]
    Yi = Y if host else [y.to(dev) for y in Y]
    res = vb.rsa_grouped(Xi, Yi, method, n_bootstrap=B, seed=7)
    idx = orc.draw_subsamples(n, B, 7)
    for i, x in enumerate(X):
        for j, y in enumerate(Y):
            ra, rb = orc.rdm(x), orc.rdm(y)
            # the device RDM differs from the CPU sgemm RDM by <= 1e-5; compare scores on the DEVICE RDMs to isolate ranking
            da, db = vb.compute_rdm(x.to(dev)).cpu(), vb.compute_rdm(y.to(dev)).cpu()
            want_pt = orc.rdm_compare(da, db, correlation=method.capitalize())
            want = orc.bootstrap_compare_with_indices(da, db, method, idx)
            got_pt, got = res["point"][i, j], res["scores"][i, j]
            for g, w in [(got_pt, want_pt)] + list(zip(got, want)):
                if math.isnan(w) or math.isnan(g):
                    assert math.isnan(w) and math.isnan(g), (trial, n, method, g, w)
                else:
                    worst = max(worst, abs(g - w))
                    assert abs(g - w) <= 1e-5, (trial, n, nA, nB, B, method, host, quant, g, w)
    print(f"trial {trial}: n={n} nA={nA} nB={nB} B={B} {method} host={host} quantised={quant} ok", flush=True)
print("fuzz ok; worst |diff| =", worst)
