"""Row-sharded execution across GPUs: nodes are split into contiguous row blocks, parameters are
replicated, and the schedule exchanges only

  C1  {S' = k^T v, z' = k^T 1, ||q||^2, ||k||^2}      one all-reduce per attention layer (h x h + 3h floats)
  C2  {dS', dz'}                                        its backward
  C3  BatchNorm column sums (forward and backward)      2h floats each
  C4  the pre-scaled SpMM operand rows                  all-gather of [N/P, h] blocks (on the reference's graphs nearly
                                                        every remote row is a halo row, so the halo IS the all-gather)
  C5  parameter gradients                               one flattened all-reduce per step

through torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests).  `Comm(None)` is the single-GPU no-op."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

Tensor = torch.Tensor


def partition(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block of `rank`: [r0, r1) with block = ceil(n / world)."""
    block = (n + world - 1) // world
    r0 = min(n, rank * block)
    return r0, min(n, r0 + block)


class Comm:
    def __init__(self, group=None, n_global: Optional[int] = None):
        self.group = group
        self.active = group is not None or (dist.is_available() and dist.is_initialized() and n_global is not None)
        if self.active:
            self.world = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
        else:
            self.world, self.rank = 1, 0
        self.n_global = n_global
        if self.active and n_global is None:
            raise ValueError("Comm needs the global node count")
        self.block = (n_global + self.world - 1) // self.world if n_global is not None else None
        self.rows = partition(n_global, self.world, self.rank) if n_global is not None else None

    # -- small reductions (C1, C2, C3, C5) ----------------------------------------------------------
    def allreduce_(self, *tensors: Tensor):
        """In-place sum over ranks; several small fp32 tensors travel as one flat buffer."""
        if not self.active or self.world == 1:
            return
        ts = [t for t in tensors if t is not None]
        if len(ts) == 1 and ts[0].is_contiguous():
            dist.all_reduce(ts[0], group=self.group)
            return
        flat = torch.cat([t.reshape(-1).float() for t in ts])
        dist.all_reduce(flat, group=self.group)
        o = 0
        for t in ts:
            k = t.numel()
            t.copy_(flat[o:o + k].view_as(t))
            o += k

    # -- C4 -----------------------------------------------------------------------------------------
    def allgather_rows(self, x_local: Tensor) -> Tensor:
        """[n_local, h] row block -> [n_global, h] (blocks of ceil(N/P) rows; the last block may be short)."""
        if not self.active or self.world == 1:
            return x_local
        n_loc, h = x_local.shape
        xl = x_local
        if n_loc != self.block or not xl.is_contiguous():
            pad = torch.zeros((self.block, h), dtype=x_local.dtype, device=x_local.device)
            pad[:n_loc] = x_local
            xl = pad
        out = torch.empty((self.world * self.block, h), dtype=x_local.dtype, device=x_local.device)
        dist.all_gather_into_tensor(out, xl, group=self.group)
        return out[:self.n_global]


SINGLE = Comm(None)
