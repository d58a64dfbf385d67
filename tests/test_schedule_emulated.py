"""Checks the forward/backward *schedules* (sgformer_b200/engine.py + functional.py) against the golden fixtures of
the reference, with the CUDA kernels replaced by their torch-CPU emulation (tests/kernel_emu.py).  Runs without a GPU.
The kernels themselves are verified on the B200 by the `-m gpu` tests."""
import glob
import os

import pytest
import torch

import kernel_emu
from sgformer_b200 import engine as E
from sgformer_b200 import functional as Fn
from sgformer_b200.config import make_config
from sgformer_b200.dist import SINGLE

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODEL_FILES = sorted(glob.glob(os.path.join(GOLD, "model_*.pt")))


@pytest.fixture(autouse=True)
def emulated_kernels(monkeypatch):
    monkeypatch.setattr(E, "K", kernel_emu)
    monkeypatch.setattr(Fn, "K", kernel_emu)
    yield


def _cfg_from_oracle(c):
    keys = make_config("large", 1, 1, 1).keys()
    kw = {k: v for k, v in c.items() if k in keys and k not in ("variant", "in_channels", "hidden", "out_channels")}
    return make_config(c["variant"], c["in_channels"], c["hidden"], c["out_channels"], **kw)


def _close(a, b, rtol, atol, what):
    a, b = a.double(), b.double()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    assert err <= atol + rtol * ref, f"{what}: max err {err:.3e} (ref max {ref:.3e})"


@pytest.mark.parametrize("path", MODEL_FILES, ids=[os.path.basename(p)[6:-3] for p in MODEL_FILES])
def test_fused_schedule_fp32(path):
    fx = torch.load(path, weights_only=False)
    cfg = _cfg_from_oracle(fx["cfg"])
    sd = {k: v.clone() for k, v in fx["state_dict"].items()}
    names = tuple(sd.keys())
    n = fx["x"].shape[0]
    graph = kernel_emu.EmuGraph(fx["edge_index"], n, 1 if cfg["variant"] == "medium" else 0)

    out = Fn.SGFormerFn.apply(fx["x"], graph, cfg, E.FP32, False, SINGLE, names, *[sd[k] for k in names])
    _close(out, fx["out_eval"], 2e-5, 2e-6, "eval output")

    params = [sd[k].clone().requires_grad_(True) if (sd[k].is_floating_point() and "running" not in k) else sd[k].clone()
              for k in names]
    x = fx["x"].clone().requires_grad_(True)
    out = Fn.SGFormerFn.apply(x, graph, cfg, E.FP32, True, SINGLE, names, *params)
    _close(out, fx["out_train"], 2e-5, 2e-6, "train output")
    (out * fx["loss_weight"]).sum().backward()
    _close(x.grad, fx["grad_x"], 5e-4, 2e-6, "grad x")
    got = dict(zip(names, params))
    for k, g in fx["grads"].items():
        assert got[k].grad is not None, f"missing grad {k}"
        _close(got[k].grad, g, 5e-4, 3e-5, f"grad {k}")
    for k, v in fx["buffers_after_train"].items():
        _close(got[k].float(), v.float(), 1e-5, 1e-6, f"buffer {k}")


def test_attention_fn_matches_reference():
    fx = torch.load(os.path.join(GOLD, "attention.pt"), weights_only=False)
    for name, c in fx.items():
        q, k, v = (c[t].clone().requires_grad_(True) for t in "qkv")
        o = Fn.AttentionFn.apply(q, k, v, E.FP32)
        _close(o, c["out"], 2e-5, 2e-6, f"{name} out")
        (o * c["w"]).sum().backward()
        for t, g in (("dq", q.grad), ("dk", k.grad), ("dv", v.grad)):
            _close(g, c[t], 1e-3, 1e-7, f"{name} {t}")


def test_bf16_schedule_is_close():
    fx = torch.load(os.path.join(GOLD, "model_large_add_init.pt"), weights_only=False)
    cfg = _cfg_from_oracle(fx["cfg"])
    sd = fx["state_dict"]
    names = tuple(sd.keys())
    graph = kernel_emu.EmuGraph(fx["edge_index"], fx["x"].shape[0], 0)
    out = Fn.SGFormerFn.apply(fx["x"], graph, cfg, E.BF16, False, SINGLE, names, *[sd[k].clone() for k in names])
    _close(out, fx["out_eval"], 3e-2, 3e-2, "bf16 eval output")
