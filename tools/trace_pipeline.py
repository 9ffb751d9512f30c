"""Timeline of rsa_grouped's copy/compute pipeline (device ms since 'start', and host ms)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import visreps_b200 as vb
from visreps_b200 import rsa, synth
dev = torch.device("cuda", 0)
feats, neur = synth.nsd_workload(1000, dev, seed=0)
Xh = [x.cpu().pin_memory() for _, x in feats]
Yh = [y.cpu().pin_memory() for _, y in neur]
import time
for rep in range(5):
    rsa.TRACE = [] if rep >= 2 else None
    t0 = time.perf_counter()
    vb.rsa_grouped(Xh, Yh, "spearman", n_bootstrap=1000)
    print(f"--- host inputs, rep {rep}: {1e3 * (time.perf_counter() - t0):.2f} ms wall")
    if rep >= 2:
        torch.cuda.synchronize()
        t0_ev, t0_host = rsa.TRACE[0][1], rsa.TRACE[0][2]
        for label, ev, th in rsa.TRACE:
            if rep == 2 or not label.startswith("copied") or label.endswith("193600)") or label.endswith("1400)"):
                print(f"{label:28s} device {t0_ev.elapsed_time(ev):8.2f} ms   host {1e3 * (th - t0_host):8.2f} ms")
rsa.TRACE = []
X = [x for _, x in feats]; Y = [y for _, y in neur]
vb.rsa_grouped(X, Y, "spearman", n_bootstrap=1000)
torch.cuda.synchronize()
t0_ev, t0_host = rsa.TRACE[0][1], rsa.TRACE[0][2]
print("--- device-resident inputs")
for label, ev, th in rsa.TRACE:
    print(f"{label:28s} device {t0_ev.elapsed_time(ev):8.2f} ms   host {1e3 * (th - t0_host):8.2f} ms")
