"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the oracle port on the host cores) prints ONE
JSON line with the keys the driver reads, and the GPU arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT)


def test_reference_arm_json_line():
    r = _run(["--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "nodes/sec fwd+bwd" and d["unit"] == "nodes/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_gpu_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return      # on a GPU box the arm runs (covered by the driver's bench run)
    r = _run(["--workload", "tiny", "--steps", "1", "--warmup", "0"], timeout=300)
    assert r.returncode != 0, "bench.py must not silently run on the CPU"
    assert not any(ln.startswith("{") for ln in r.stdout.splitlines())
