"""Multi-GPU parity: the ladder sharded over 2, 4 and 8 GPUs gives the identical
result to the single-process oracle — the reference's invariance-to-#workers
guarantee (docs/src/distributed.md:37-55, test/test_parallelism_invariance.jl)."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_ladder_matches_oracle(world, tmp_path):
    if n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    out = tmp_path / "report.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "multigpu_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    report = json.loads(out.read_text())
    assert all(v == [] for v in report.values()), report
