"""Row-sharded execution over NCCL on real GPUs == the single-GPU result (logits and every parameter gradient, biases
included).  Needs >= 2 GPUs in the box: skipped on the single-GPU test box; the same schedule is covered on CPU by
tests/test_row_sharding_gloo.py (gloo, world 2 and 3)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4])
def test_row_sharded_nccl_matches_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, the box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    info = "\n".join(ln for ln in (r.stdout + "\n" + r.stderr).splitlines()
                     if any(t in ln for t in ("rel err", "multi_gpu_check", "Error", "error", "Traceback", "File \"/root")))[-4000:]
    assert "multi_gpu_check: OK" in r.stdout, info
    assert r.returncode == 0, "checks passed but a rank exited with an error (teardown):\n" + info
