"""Stage breakdown of the THINGS-shape job (n = 1854) on one GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import visreps_b200 as vb
from visreps_b200 import synth, _lib
dev = torch.device("cuda", 0)
n = int(os.environ.get("QN", 1854))
z = synth.latent(n, dev, 1)
feats = [synth.layer_features(z, d, 10 + i) for i, d in enumerate(synth.ALEXNET_DIMS.values())]
g = torch.Generator(device=dev).manual_seed(7)
beh = torch.relu(z @ torch.randn(synth.LATENT, 66, device=dev, generator=g))
_lib.set_option('sweep_tile_a', int(os.environ.get('TA', 8))); _lib.set_option('sweep_tile_b', int(os.environ.get('TB', 8)))
for rep in range(3):
    if rep == 2: _lib.profile_start()
    torch.cuda.synchronize(); t = time.perf_counter()
    res = vb.rsa_grouped(feats, [beh], "spearman", n_bootstrap=1000)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"n={n}: {dt*1e3:.1f} ms; stages:", {k: round(v["ms"], 2) for k, v in _lib.profile_stop().items() if v["ms"] > 0})
