"""NCCL coverage of visreps_b200.dist: the sharded RDM build + blob exchange + replicate sharding must be
bit-identical to the single-GPU job.  One process per GPU through torchrun; the one-rank case runs on any GPU box,
the two-rank case needs two GPUs (`gpurun --gpus 2`)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

import visreps_b200 as vb
from visreps_b200 import _lib, rsa as vrsa

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "_nccl_sharded_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert f"nccl sharded build ok: world={world}" in p.stdout


def test_sharded_build_one_rank_nccl():
    _run(1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_build_two_ranks_nccl():
    _run(2)


def test_presorted_blob_round_trip():
    """A blob copied elsewhere (what the all-gather delivers) and adopted with from_blob ranks exactly like the original."""
    dev = torch.device("cuda", 0)
    rs = np.random.RandomState(3)
    n = 130
    a = vb.compute_rdm(torch.from_numpy(rs.standard_normal((n, 40)).astype(np.float32)).to(dev))
    b = vb.compute_rdm(torch.from_numpy(np.round(rs.standard_normal((n, 6)), 1).astype(np.float32)).to(dev))   # ties
    pa, pb = vrsa.PresortedRDM(a), vrsa.PresortedRDM(b)
    lib = _lib.load()
    stacked = torch.stack([pa.blob, pb.blob]).clone()
    assert stacked.shape[1] == lib.vr_presort_bytes(n)
    qa, qb = vrsa.PresortedRDM.from_blob(stacked[0], n), vrsa.PresortedRDM.from_blob(stacked[1], n)
    masks = vrsa.build_masks(vrsa.draw_index_sets(n, 9, 42), n, dev, with_full=True)
    want = vrsa.compare_grouped([pa], [pb], "spearman", masks)
    got = vrsa.compare_grouped([qa], [qb], "spearman", masks)
    assert torch.equal(want, got)
    with pytest.raises(ValueError):
        vrsa.PresortedRDM.from_blob(stacked[0][:-8], n)
    with pytest.raises(ValueError):
        vrsa.PresortedRDM.from_blob(stacked[0].cpu(), n)
