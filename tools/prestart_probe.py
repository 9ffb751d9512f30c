"""Where does the host spend time between entering rsa_grouped and the first enqueue? (host-input path)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import visreps_b200 as vb
from visreps_b200 import rsa, synth, _lib
dev = torch.device("cuda", 0)
feats, neur = synth.nsd_workload(1000, dev, seed=0)
Xh = [x.cpu().pin_memory() for _, x in feats]
Yh = [y.cpu().pin_memory() for _, y in neur]
lib = _lib.load()
def T(label, fn):
    t = time.perf_counter(); r = fn(); print(f"   {label:34s} {1e3 * (time.perf_counter() - t):8.3f} ms"); return r
for rep in range(4):
    t = time.perf_counter()
    vb.rsa_grouped(Xh, Yh, "spearman", n_bootstrap=1000)
    print(f"rep {rep}: {1e3 * (time.perf_counter() - t):.2f} ms")
    T("mem_get_info", lambda: torch.cuda.mem_get_info(dev))
    T("memory_reserved", lambda: torch.cuda.memory_reserved(dev))
    T("memory_allocated", lambda: torch.cuda.memory_allocated(dev))
    T("workspace", lambda: _lib.workspace(max(lib.vr_rdm_workspace_bytes(1000, int(x.shape[1])) for x in Xh + Yh), dev))
    T("pinned empty 512K", lambda: torch.empty((8, 8, 1001), dtype=torch.float64, pin_memory=True))
    T("pinned empty 128K", lambda: torch.empty((1000, 32), dtype=torch.int32, pin_memory=True))
    T("device alloc 5.6MB", lambda: torch.empty(1400 * 1000, device=dev))
    T("device alloc 774MB", lambda: torch.empty(193600 * 1000, device=dev))
    T("mem_get_info again", lambda: torch.cuda.mem_get_info(dev))
    import gc
    T("gc.collect", gc.collect)
print(torch.cuda.memory_summary(dev, abbreviated=True)[:1500])
