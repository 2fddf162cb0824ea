"""Two-GPU checks (skipped on a single-GPU box): the data-parallel update through the fused peer-memory
gradient-sum + Adam kernel must leave bit-identical parameters on both ranks and match the NCCL-allreduce path."""
import os
import re
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "dist_check.py")] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_update_p2p_matches_nccl():
    a = _run([], 29561)
    b = _run(["--nccl"], 29562)
    da = re.findall(r"param digest ([0-9a-f]{32})", a)
    db = re.findall(r"param digest ([0-9a-f]{32})", b)
    assert len(da) == 2 and len(set(da)) == 1, da          # both ranks hold identical parameters
    assert set(da) == set(db), (da, db)                    # fused peer-memory path == NCCL allreduce path (2 ranks: a+b)
    assert "p2p adam attached: True" in a and "p2p adam attached: False" in b
    assert "p2p timed out: False" in a
