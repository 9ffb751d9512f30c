"""Pearson bootstrap timing at the NSD shape: 8 x 8 RDM pairs x 1001 masks."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import visreps_b200 as vb
from visreps_b200 import rsa as vrsa, synth
dev = torch.device("cuda", 0)
n = 1000
z = synth.latent(n, dev, 1)
A = [vrsa._RawRDM(vb.compute_rdm(synth.layer_features(z, 2048, 10 + i))) for i in range(8)]
B = [vrsa._RawRDM(vb.compute_rdm(z @ torch.randn(synth.LATENT, 300 + 50 * i, device=dev) + 0.5 * torch.randn(n, 300 + 50 * i, device=dev))) for i in range(8)]
masks = vrsa.build_masks(vrsa.draw_index_sets(n, 1000, 42), n, dev, with_full=True)
for ka, kb in ((1, 1), (8, 8)):
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        out = vrsa.compare_grouped(A[:ka], B[:kb], "pearson", masks)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"pearson bootstrap {ka}x{kb}x{masks.shape[0]} n={n}: {dt*1e3:.3f} ms -> {ka*kb*masks.shape[0]/dt:.0f} replicate-compares/s", flush=True)
