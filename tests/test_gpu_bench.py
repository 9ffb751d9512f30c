"""bench.py's own arm on a GPU at a shrunken size: the process exits 0 and prints ONE JSON line with every key of the contract
(this is the command the driver runs; a crash here is a lost round-end number)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_own_arm_prints_one_contract_line_small():
    env = dict(os.environ, VR_BENCH_NO_CLOCKS="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--small", "--n", "200", "--n-bootstrap", "40", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [l for l in p.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    gates = d["cpu_baseline"].get("parity_gates")
    assert gates and gates["rdm_ok"] and gates["ranks_bit_exact"] and gates["scores_ok"], gates
