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

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist

Tensor = torch.Tensor

# Column chunks of the C4 all-gather / SpMM pipeline (1 = one blocking all-gather per SpMM).  Measured on 2 B200s at the
# products shape (profiles/r1_scaling.md): 60.1 ms/step with 1 chunk, 61.2 with 2, 61.3 with 4 - NCCL's all-gather kernels need SMs
# that the HBM-bound SpMM grid already fills, so the transfer is not hidden and the narrower SpMMs cost more; default stays 1.
C4_CHUNKS = int(os.environ.get("SGF_C4_CHUNKS", "1"))
C4_MIN_CHUNK_BYTES = 256      # gathered rows stay >= two full 128-byte lines per neighbour


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

    def c4_chunks(self, h: int, elem_size: int) -> int:
        """Column chunks of the C4 pipeline (C4_CHUNKS, reduced until the chunk rows are >= C4_MIN_CHUNK_BYTES)."""
        if not self.active or self.world == 1:
            return 1
        want = C4_CHUNKS
        while want > 1 and (h % want != 0 or (h // want) * elem_size < C4_MIN_CHUNK_BYTES):
            want -= 1
        return want

    def spmm_gathered(self, spmm, rowptr: Tensor, col: Tensor, row_scale: Optional[Tensor], x_local: Tensor, heavy=None) -> Tensor:
        """C4 + SpMM: y[rows of this rank] = scale * A[rows, :] @ all_gather(x_local).

        The operand is exchanged in column chunks, each an asynchronous all-gather on NCCL's stream; the SpMM of chunk c
        (`spmm(rowptr, col, row_scale, x_chunk, out=y[:, chunk], heavy=...)`, i.e. kernels.spmm) starts as soon as chunk c
        has arrived and overlaps the transfer of chunk c+1.  On uniform random graphs every remote row is a halo row, so
        the exchange cannot be smaller than the all-gather; it can only be hidden."""
        if not self.active or self.world == 1:
            return spmm(rowptr, col, row_scale, x_local, heavy=heavy)
        n_loc, h = x_local.shape
        nch = self.c4_chunks(h, x_local.element_size())
        if nch == 1:
            return spmm(rowptr, col, row_scale, self.allgather_rows(x_local), heavy=heavy)
        hc = h // nch
        dev, dt = x_local.device, x_local.dtype
        pending = []
        for c in range(nch):
            xc = torch.zeros((self.block, hc), dtype=dt, device=dev) if n_loc != self.block else \
                torch.empty((self.block, hc), dtype=dt, device=dev)
            xc[:n_loc].copy_(x_local[:, c * hc:(c + 1) * hc])
            gc = torch.empty((self.world * self.block, hc), dtype=dt, device=dev)
            pending.append((dist.all_gather_into_tensor(gc, xc, group=self.group, async_op=True), gc, xc))
        n_rows = rowptr.numel() - 1
        # same pitch rule as kernels.alloc_act (rows 16-byte aligned); chunk offsets are multiples of 256 bytes
        out = torch.empty((n_rows, h), dtype=dt, device=dev)
        for c, (work, gc, _) in enumerate(pending):
            work.wait()
            spmm(rowptr, col, row_scale, gc[:self.n_global], out=out[:, c * hc:(c + 1) * hc], heavy=heavy)
        return out


SINGLE = Comm(None)
