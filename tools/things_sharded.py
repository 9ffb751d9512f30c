"""BASELINE config 3 under torchrun: THINGS shape (n = 1854, behavioural RDM d = 66 vs 8 AlexNet layers), 1000
replicates sharded over the ranks (visreps_b200.dist.rsa_grouped_sharded, one NCCL all-gather of the scores).
Rank 0 also runs the job alone and checks that the gathered result is identical."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import visreps_b200 as vb
from visreps_b200 import synth, dist as vdist, rsa as vrsa
vrsa.MASK_CACHE = os.environ.get("MASK_CACHE", "1") != "0"       # 0: every call replays the index draws (as bench.py does)
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n, B = int(os.environ.get("QN", 1854)), int(os.environ.get("QB", 1000))
z = synth.latent(n, dev, 1)                                   # same seeds on every rank: identical inputs
feats = [synth.layer_features(z, d, 10 + i) for i, d in enumerate(synth.ALEXNET_DIMS.values())]
g = torch.Generator(device=dev).manual_seed(7)
beh = torch.relu(z @ torch.randn(synth.LATENT, 66, device=dev, generator=g))
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
for rep in range(3):
    barrier(); t = time.perf_counter()
    res = vdist.rsa_grouped_sharded(feats, [beh], "spearman", n_bootstrap=B)
    barrier(); dt = time.perf_counter() - t
    if rank == 0:
        print(f"THINGS shape n={n}, {world} GPU(s): 8 layers x 1 behavioural RDM x {B} replicates: {dt*1e3:.1f} ms -> {8*B/dt:.0f} replicate-compares/s", flush=True)
if rank == 0:
    alone = vb.rsa_grouped(feats, [beh], "spearman", n_bootstrap=B)
    diff = max(float(np.abs(alone["scores"] - res["scores"]).max()), float(np.abs(alone["point"] - res["point"]).max()))
    print(f"max |sharded (K-split RDM builds) - single GPU| over all scores: {diff:.2e} | point", np.round(res["point"][:, 0], 4))
    assert diff <= 1e-5
if world > 1:
    exact = vdist.rsa_grouped_sharded(feats, [beh], "spearman", n_bootstrap=B, ksplit=False)
    if rank == 0:
        same = np.array_equal(alone["scores"], exact["scores"]) and np.array_equal(alone["point"], exact["point"]) and np.array_equal(alone["ci_low"], exact["ci_low"])
        print("sharded with ksplit=False == single-GPU result, bit for bit:", same)
        assert same
if world > 1:
    dist.barrier(); dist.destroy_process_group()
