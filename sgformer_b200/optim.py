"""Fused two-group Adam on the device.

Drop-in for the reference's `torch.optim.Adam([{'params': model.params1, 'weight_decay': a}, {'params': model.params2,
'weight_decay': b}], lr=lr)` (large/main.py:115-119, medium/main.py:112-117): same constructor, `step()`, `zero_grad()`,
`param_groups`, `state_dict()` layout (per-parameter `step`, `exp_avg`, `exp_avg_sq`), but every parameter tensor of a step is
updated by one kernel launch (`sgf_adam_step`, csrc/optim.cu) and the step counts are device scalars, so the whole training step
can be captured in a CUDA graph.  CUDA fp32 parameters only; there is no CPU path."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import SGF_ADAM_MAX_TENSORS, AdamArgs, check


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))

    def _state(self, p: torch.Tensor):
        st = self.state[p]
        if not st:
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)      # device scalar, advanced by the kernel
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        elif st["step"].device != p.device:       # parameters were moved after construction (model.to(device), main-batch.py:131)
            for k_ in ("step", "exp_avg", "exp_avg_sq"):
                st[k_] = st[k_].to(p.device)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        items = []
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise RuntimeError("sgformer_b200.optim.Adam updates CUDA fp32 parameters only (no CPU fallback)")
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("sgformer_b200.optim.Adam needs contiguous parameters and gradients")
                st = self._state(p)
                items.append((p, p.grad, st["exp_avg"], st["exp_avg_sq"], float(group["lr"]), b1, b2, float(group["eps"]),
                              float(group["weight_decay"]), st["step"]))
        if not items:
            return loss
        dev = items[0][0].device
        if any(it[0].device != dev for it in items):
            raise RuntimeError("sgformer_b200.optim.Adam: all parameters must live on one device")
        with torch.cuda.device(dev):
            check(lib.sgf_set_device(dev.index if dev.index is not None else torch.cuda.current_device()), "sgf_set_device")
            stream = torch.cuda.current_stream().cuda_stream
            for o in range(0, len(items), SGF_ADAM_MAX_TENSORS):
                part = items[o:o + SGF_ADAM_MAX_TENSORS]
                a = AdamArgs()
                a.n_tensors = len(part)
                for i, (p, g, m, v, lr, b1, b2, eps, wd, stp) in enumerate(part):
                    a.param[i], a.grad[i], a.exp_avg[i], a.exp_avg_sq[i] = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
                    a.numel[i] = p.numel()
                    a.lr[i], a.beta1[i], a.beta2[i], a.eps[i], a.weight_decay[i] = lr, b1, b2, eps, wd
                    a.step[i] = stp.data_ptr()
                check(lib.sgf_adam_step(C.byref(a), stream), "sgf_adam_step")
        return loss
