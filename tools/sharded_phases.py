"""Phase costs of dist.rsa_grouped_sharded at the THINGS shape (synchronised between phases, so the sum exceeds the overlapped
job): where the strong-scaling time goes.  torchrun --nproc-per-node N tools/sharded_phases.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import visreps_b200 as vb
from visreps_b200 import synth, dist as vdist, rsa as vrsa, _lib
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n, B = 1854, 1000
z = synth.latent(n, dev, 1)
feats = [synth.layer_features(z, d, 10 + i) for i, d in enumerate(synth.ALEXNET_DIMS.values())]
beh = torch.relu(z @ torch.randn(synth.LATENT, 66, device=dev, generator=torch.Generator(device=dev).manual_seed(7)))
def sync():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
for rep in range(3):
    vdist.rsa_grouped_sharded(feats, [beh], "spearman", n_bootstrap=B)
marks = []
def mark(label):
    sync(); marks.append((label, time.perf_counter()))
for rep in range(2):
    marks.clear(); mark("start")
    inputs = feats + [beh]
    rdms = []
    for x in inputs:
        if x.shape[1] >= vdist.KSPLIT_MIN_COLS * world and world > 1:
            lo, hi = vdist.column_shard(x.shape[1], rank, world)
            rdms.append(vdist.compute_rdm_ksplit(x[:, lo:hi], d_total=x.shape[1]))
        else:
            rdms.append(vb.compute_rdm(x))
    mark("rdm builds (K-split where wide; narrow ones on every rank here)")
    mine = vrsa.PresortedRDM.batch(rdms[rank::world]) if rdms[rank::world] else []
    mark("presort of this rank's share")
    lo, hi = vdist.shard_bounds(B, rank, world)
    t0 = time.perf_counter(); draws = vrsa.draw_subsample_masks(n, B, 42, replicate_range=(lo, hi)); t_draw = time.perf_counter() - t0
    masks = vrsa.build_masks(draws, n, dev, with_full=True)
    mark(f"host draw ({t_draw*1e3:.2f} ms host) + masks")
    items = vrsa.PresortedRDM.batch(rdms)
    mark("(presort of all 9, only to have blobs for the next phase)")
    _lib.profile_start()
    out = vrsa.compare_grouped(items[:8], items[8:], "spearman", masks)
    mark("count tables + tie sums + sweep of this rank's replicates")
    prof = _lib.profile_stop()
    if rank == 0:
        for (a, ta), (b, tb) in zip(marks, marks[1:]):
            print(f"  {b:75s} {1e3*(tb-ta):8.2f} ms")
        print("  stages:", {k: round(v['ms'], 2) for k, v in prof.items() if v['ms'] > 0})
sync()
for rep in range(3):
    sync(); t0 = time.perf_counter(); vdist.rsa_grouped_sharded(feats, [beh], "spearman", n_bootstrap=B); sync()
    if rank == 0: print(f"whole job on {world} GPU(s): {(time.perf_counter()-t0)*1e3:.2f} ms")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
