"""BASELINE config 3 timing: THINGS shape, 1854-stimulus behavioural RDM (d=66) vs 8 AlexNet layers, 1000 replicates."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import visreps_b200 as vb
from visreps_b200 import synth
dev = torch.device("cuda", 0)
n = 1854
z = synth.latent(n, dev, 1)
feats = [synth.layer_features(z, d, 10 + i) for i, d in enumerate(synth.ALEXNET_DIMS.values())]
beh = torch.relu(z @ torch.randn(synth.LATENT, 66, device=dev, generator=torch.Generator(device=dev).manual_seed(7)))
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    res = vb.rsa_grouped(feats, [beh], "spearman", n_bootstrap=1000)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"THINGS shape n={n}: 8 layers x 1 behavioural RDM x 1000 replicates: {dt*1e3:.1f} ms -> {8*1000/dt:.0f} replicate-compares/s", flush=True)
print("point", res["point"][:, 0])
