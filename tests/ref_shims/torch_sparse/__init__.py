"""Shim of torch_sparse==0.6.10 (test infra): SparseTensor + matmul (sum reduce).

Published semantics restated: SparseTensor(row, col, value, sparse_sizes) sorts entries by
(row, col) keeping duplicates; matmul(A, X) = CSR SpMM with sum reduction, differentiable in X
(and in value). Reference call sites: large/ours.py:33-34, 100M/ours.py:79-80."""
import torch


class SparseTensor:
    def __init__(self, row=None, rowptr=None, col=None, value=None, sparse_sizes=None,
                 is_sorted=False, trust_data=False):
        n_rows, n_cols = sparse_sizes
        if not is_sorted:
            key = row.to(torch.int64) * n_cols + col.to(torch.int64)
            perm = torch.argsort(key, stable=True)
            row, col = row[perm], col[perm]
            if value is not None:
                value = value[perm]
        self._row, self._col, self._value = row, col, value
        self._sizes = (int(n_rows), int(n_cols))
        counts = torch.bincount(row, minlength=n_rows)
        self._rowptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=row.device)
        self._rowptr[1:] = torch.cumsum(counts, 0)

    def sparse_sizes(self):
        return self._sizes

    def csr(self):
        return self._rowptr, self._col, self._value

    def coo(self):
        return self._row, self._col, self._value

    def to(self, *a, **k):
        return self


def matmul(src, other, reduce="sum"):
    assert reduce in ("sum", "add")
    row, col, val = src.coo()
    msgs = other[col]
    if val is not None:
        msgs = msgs * val.unsqueeze(-1).to(other.dtype)
    out = torch.zeros((src.sparse_sizes()[0],) + tuple(other.shape[1:]), dtype=other.dtype,
                      device=other.device)
    return out.index_add(0, row, msgs)
