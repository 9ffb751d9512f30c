"""One NSD-shape grouped Spearman job from device-resident RDMs (for ncu: count / tie / sweep kernels only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import visreps_b200 as vb
from visreps_b200 import rsa as vrsa, synth
dev = torch.device("cuda", 0)
n = int(os.environ.get("QN", 1000)); B = int(os.environ.get("QB", 1000))
z = synth.latent(n, dev, 1)
feats = [synth.layer_features(z, d, 10 + i) for i, d in enumerate([4096, 4096, 1000, 2048, 3000, 1500, 2500, 3500])]
rois = [synth.roi_responses(z, v, 500 + i) for i, v in enumerate(synth.NSD_ROIS.values())]
A = vrsa.PresortedRDM.batch([vb.compute_rdm(x) for x in feats])
Bn = vrsa.PresortedRDM.batch([vb.compute_rdm(y) for y in rois])
masks = vrsa.build_masks(vrsa.draw_subsample_masks(n, B, 42), n, dev, with_full=True)
from visreps_b200 import _lib
_lib.set_option('sweep_tile_a', int(os.environ.get('TA', 8))); _lib.set_option('sweep_tile_b', int(os.environ.get('TB', 8)))
for rep in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = vrsa.compare_grouped(A, Bn, "spearman", masks)
    e1.record(); torch.cuda.synchronize()
    print(f"grouped 8x8x{B + 1}: {e0.elapsed_time(e1):.2f} ms", float(out[0, 0, 0]))
