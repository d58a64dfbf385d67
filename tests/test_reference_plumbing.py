"""BASELINE config 1 (plumbing, no GPU): the reference's own `medium/main.py --method ours --cpu` runs end to end in the
build container on a Cora-shaped synthetic dataset through tests/ref_shims — proves the shims + driver wiring the golden
fixtures rely on.  (On the GPU box the same driver would run with the drop-in `ours.py`; /root/reference does not exist
there, so this test is build-container only.)"""
import os
import subprocess
import sys

import pytest

from _refload import REF_ROOT, SHIMS, reference_available


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container)")
def test_medium_main_runs_on_cora_shaped_synthetic(tmp_path):
    env = dict(os.environ, PYTHONPATH=SHIMS, OMP_NUM_THREADS="4")
    cmd = [sys.executable, os.path.join(REF_ROOT, "medium", "main.py"), "--backbone", "gcn", "--dataset", "cora", "--lr", "0.01",
           "--num_layers", "4", "--hidden_channels", "64", "--weight_decay", "5e-4", "--dropout", "0.5", "--method", "ours",
           "--ours_layers", "1", "--use_graph", "--graph_weight", "0.8", "--ours_dropout", "0.2", "--use_residual", "--alpha",
           "0.5", "--ours_weight_decay", "0.001", "--rand_split_class", "--valid_num", "500", "--test_num", "1000",
           "--no_feat_norm", "--seed", "123", "--cpu", "--runs", "1", "--epochs", "3", "--data_dir", str(tmp_path) + "/"]
    r = subprocess.run(cmd, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Epoch: 02" in r.stdout or "Epoch: 00" in r.stdout
    assert os.path.exists(os.path.join(str(tmp_path), "results", "cora_ours_gcn.txt"))
