"""Fused training loss: log_softmax + NLL over the training rows in one kernel, forward and
logits-gradient together.  Replaces `F.log_softmax(out, 1)` + `nn.NLLLoss()(out[train_mask], y[train_mask])`
(reference large/main.py:139-141, large/main-batch.py:147-148)."""
from typing import Optional

import torch

from . import functional as Fn


def nll_loss_from_logits(logits: torch.Tensor, labels: torch.Tensor, mask: Optional[torch.Tensor] = None,
                         denom: Optional[float] = None) -> torch.Tensor:
    """Mean over the rows selected by `mask` (bool [N]; all rows when None).  `denom` overrides the divisor (e.g. the global
    node count in a row-sharded run); when omitted with a mask it is mask.sum() (one device sync — pass it for speed)."""
    if not logits.is_cuda:
        raise RuntimeError("sgformer_b200.loss needs CUDA tensors (no CPU fallback)")
    if denom is None:
        denom = float(mask.sum().item()) if mask is not None else float(logits.shape[0])
    # the kernel takes a row pitch: the encoder's padded-pitch logits ([N, 47] inside [N, 48]) are read in place
    ok = logits.dtype == torch.float32 and logits.dim() == 2 and logits.stride(1) == 1
    return Fn.SoftmaxNLLFn.apply(logits if ok else logits.float().contiguous(), labels, mask, denom)
