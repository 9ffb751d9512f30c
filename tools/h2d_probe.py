"""Concurrent pinned host->device bandwidth per rank, with and without NUMA-local pinned buffers (torchrun)."""
import os, sys, time, glob
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
def numa_of_gpu(i):
    try:
        bus = torch.cuda.get_device_properties(i).pci_bus_id
        dom = torch.cuda.get_device_properties(i).pci_domain_id
        devid = torch.cuda.get_device_properties(i).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        return int(open(path).read().strip()), path
    except Exception as e:
        return -1, repr(e)
def cpus_of_node(node):
    try:
        txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        out = []
        for part in txt.split(","):
            a, _, b = part.partition("-")
            out += list(range(int(a), int(b or a) + 1))
        return out
    except Exception:
        return []
def bw(label):
    x = torch.empty(256 << 20, dtype=torch.float32).pin_memory()      # 1 GiB
    x.fill_(1.0)
    d = torch.empty_like(x, device=dev)
    for _ in range(2): d.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(4): d.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print(f"rank {rank} {label}: {4 * x.numel() * 4 / dt / 1e9:.1f} GB/s  affinity={len(os.sched_getaffinity(0))} cpus", flush=True)
    del x, d
node, info = numa_of_gpu(lr)
print(f"rank {rank}: gpu numa node {node} ({info}); nodes present: {sorted(glob.glob('/sys/devices/system/node/node*'))[:4]} cpus={os.cpu_count()}", flush=True)
bw("default")
cp = cpus_of_node(node) if node >= 0 else []
if cp:
    os.sched_setaffinity(0, set(cp) & os.sched_getaffinity(0) or os.sched_getaffinity(0))
    bw("numa-local")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
