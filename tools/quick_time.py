"""Ad-hoc device timings of the individual stages (development aid, not the bench contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import visreps_b200 as vb
from visreps_b200 import rsa as vrsa, _lib


def ev_time(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    dev = torch.device("cuda", 0)
    n = int(os.environ.get("QN", 1000))
    B = int(os.environ.get("QB", 1000))
    nA = int(os.environ.get("QA", 8))
    g = torch.Generator(device=dev).manual_seed(0)
    z = torch.randn(n, 64, device=dev, generator=g)
    dims = [4096, 1000, 43264][: max(1, min(3, nA))]
    for d in dims + [193600]:
        x = torch.relu(z @ (torch.randn(64, d, device=dev, generator=g) / 8) + torch.randn(n, d, device=dev, generator=g))
        t = ev_time(lambda: vb.compute_rdm(x))
        r = vb.compute_rdm(x)
        xd = x.double(); xd = xd - xd.mean(1, keepdim=True); sd = (xd.pow(2).mean(1) + 1e-12).sqrt()
        truth = 1 - ((xd @ xd.T) / d / (sd[:, None] * sd[None, :] + 1e-12)).clamp(-1, 1); truth.fill_diagonal_(0)
        err = float((r.double() - truth).abs().max())
        del xd, truth
        print(f"compute_rdm n={n} d={d}: {t:.3f} ms  -> {n*(n+1)*d/t/1e9:.2f} TFLOP/s (n(n+1)d)  max|err| vs fp64 = {err:.2e}", flush=True)
        del x
    models = [vb.compute_rdm(torch.relu(z @ (torch.randn(64, 512, device=dev, generator=g) / 8) + torch.randn(n, 512, device=dev, generator=g))) for _ in range(nA)]
    neurals = [vb.compute_rdm(z @ torch.randn(64, 300, device=dev, generator=g) + 0.5 * torch.randn(n, 300, device=dev, generator=g)) for _ in range(nA)]
    t = ev_time(lambda: vrsa.PresortedRDM(models[0]))
    print(f"presort n={n}: {t:.3f} ms", flush=True)
    A = [vrsa.PresortedRDM(m) for m in models]
    Bn = [vrsa.PresortedRDM(m) for m in neurals]
    idx = vrsa.draw_index_sets(n, B, 42)
    masks = vrsa.build_masks(idx, n, dev, with_full=True)
    for (ka, kb) in [(1, 1), (nA, nA)]:
        fn = lambda: vrsa.compare_grouped(A[:ka], Bn[:kb], "spearman", masks)
        t = ev_time(fn, iters=2)
        reps = ka * kb * masks.shape[0]
        print(f"spearman bootstrap {ka}x{kb}x{masks.shape[0]} n={n}: {t:.3f} ms -> {reps/t*1e3:.0f} replicate-compares/s", flush=True)
    fn = lambda: vrsa.compare_grouped([vrsa._RawRDM(models[0])], [vrsa._RawRDM(neurals[0])], "pearson", masks)
    t = ev_time(fn, iters=2)
    print(f"pearson bootstrap 1x1x{masks.shape[0]}: {t:.3f} ms", flush=True)
    print("launches", _lib.launch_count())


if __name__ == "__main__":
    main()
