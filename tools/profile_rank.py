"""One warm + one profiled rsa_grouped call at the NSD shape (device-resident inputs); meant to run under
ncu with `-k regex:rank_kernel --launch-skip 4 --launch-count 4` (each call launches the tie-free and the
general sweep for the materialised side and for the fused side)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import visreps_b200 as vb
from visreps_b200 import synth
dev = torch.device("cuda", 0)
feats, neur = synth.nsd_workload(int(os.environ.get("QN", 1000)), dev, seed=0)
X = [x for _, x in feats]; Y = [y for _, y in neur]
for _ in range(2):
    res = vb.rsa_grouped(X, Y, "spearman", n_bootstrap=int(os.environ.get("QB", 1000)))
torch.cuda.synchronize()
print("point[0,0]", res["point"][0, 0])
