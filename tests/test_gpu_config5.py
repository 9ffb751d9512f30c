"""BASELINE config 5 at its full size (n = 10,000 stimuli x 193,600 features, N = 49,995,000 RDM entries) through the
size-independent properties of tools/stress_check.py: 64 RDM rows against the same formula in fp64, exact symmetry, a
self-comparison of 1, and a bootstrap replicate equal to the direct comparison of the gathered sub-RDMs.  ~25 GB, ~20 s."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_config5_full_size_properties():
    free, _total = torch.cuda.mem_get_info(0)
    if free < 60 * (1 << 30):
        pytest.skip("needs ~25 GB of free device memory plus headroom")
    env = dict(os.environ, SB="40")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_check.py")], capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-2000:])
    assert "stress check ok" in p.stdout
    m = re.search(r"RDM max-abs error vs fp64 on 64 rows x 10000 columns: ([0-9.e+-]+)", p.stdout)
    assert m and float(m.group(1)) <= 1e-5, p.stdout[-2000:]
    m = re.search(r"replicate 0 vs direct sub-RDM comparison: ([0-9.e+-]+) ([0-9.e+-]+)", p.stdout)
    assert m and abs(float(m.group(1)) - float(m.group(2))) <= 1e-12
