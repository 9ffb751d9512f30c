"""BASELINE config 5 RDM with the features split over the ranks (K-split SYRK + NCCL all-reduce of the Gram).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/ksplit_rdm.py
Every rank builds the same synthetic (n, d) matrix, keeps its column shard, and the ranks cooperate; rank 0 also runs the
single-GPU RDM and checks 64 rows against fp64.  KN / KD env: n, d."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import visreps_b200 as vb
from visreps_b200 import synth, dist as vdist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n, d = int(os.environ.get("KN", 10000)), int(os.environ.get("KD", 193600))
x = synth.layer_features(synth.latent(n, dev, 1), d, 11)
lo, hi = vdist.column_shard(d, rank, world)
shard = x[:, lo:hi].contiguous()
def sync():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
for rep in range(3):
    sync(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); t0 = time.perf_counter(); e0.record()
    rdm = vdist.compute_rdm_ksplit(shard, d_total=d)
    e1.record(); sync(); wall = time.perf_counter() - t0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"K-split RDM n={n} d={d} on {world} GPU(s): {float(ms):.2f} ms device (max over ranks), {wall*1e3:.2f} ms wall -> "
              f"{n*(n+1)*d/float(ms)/1e9:.1f} TFLOP/s algorithmic", flush=True)
if rank == 0:
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); one = vb.compute_rdm(x); torch.cuda.synchronize()
        print(f"single-GPU compute_rdm: {(time.perf_counter()-t0)*1e3:.2f} ms", flush=True)
    print("max |K-split - single GPU| =", float((rdm - one).abs().max()))
    rows = torch.linspace(0, n - 1, 64, device=dev).long()
    s1 = torch.zeros(n, dtype=torch.float64, device=dev); s2 = torch.zeros_like(s1); g = torch.zeros((64, n), dtype=torch.float64, device=dev)
    for c0 in range(0, d, 16384):
        blk = x[:, c0:c0 + 16384].double(); s1 += blk.sum(1); s2 += (blk * blk).sum(1); g += blk[rows] @ blk.T
    mean = s1 / d; cov = g / d - mean[rows][:, None] * mean[None, :]
    sd = (s2 / d - mean * mean).clamp_(min=0).add(1e-12).sqrt()
    truth = 1.0 - (cov / (sd[rows][:, None] * sd[None, :] + 1e-12)).clamp_(-1, 1)
    truth[torch.arange(64, device=dev), rows] = 0.0
    err = float((rdm[rows].double() - truth).abs().max())
    print(f"K-split RDM max-abs error vs fp64 on 64 rows: {err:.3e}")
    assert err <= 1e-5 and torch.equal(rdm, rdm.T)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
