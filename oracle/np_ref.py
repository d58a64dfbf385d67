"""Independent fp64 numpy/scipy restatements of the two core contractions — TEST INFRASTRUCTURE ONLY.

Used to cross-check `oracle/sgformer_oracle.py` (and through it the CUDA kernels) against arithmetic
that shares no code with torch: plain einsum for the linear attention (medium/ours.py:14-34) and
scipy CSR for Â·X (large/ours.py:25-34).  See the header of sgformer_oracle.py for who may import oracle/.
"""
import numpy as np
import scipy.sparse as sp


def attention_fp64(q, k, v):
    """q,k [N,H,M], v [N,H,D] -> dict(out [N,H,D], S' [H,M,D], z' [H,M], nq2, nk2) in float64."""
    q, k, v = (np.asarray(t, dtype=np.float64) for t in (q, k, v))
    n = q.shape[0]
    nq2, nk2 = float((q * q).sum()), float((k * k).sum())
    s_raw = np.einsum("lhm,lhd->hmd", k, v)
    z_raw = k.sum(axis=0)
    qn = q / np.sqrt(nq2)
    num = np.einsum("nhm,hmd->nhd", qn, s_raw / np.sqrt(nk2)) + n * v
    den = np.einsum("nhm,hm->nh", qn, z_raw / np.sqrt(nk2)) + n
    return dict(out=num / den[..., None], S=s_raw, z=z_raw, nq2=nq2, nk2=nk2)


def attention_grads_fp64(q, k, v, g):
    """Closed-form backward of `attention_fp64`; g = dL/dout [N,H,D]."""
    q, k, v, g = (np.asarray(t, dtype=np.float64) for t in (q, k, v, g))
    n = q.shape[0]
    nq, nk = np.sqrt((q * q).sum()), np.sqrt((k * k).sum())
    qn, kn = q / nq, k / nk
    s = np.einsum("lhm,lhd->hmd", kn, v)
    z = kn.sum(axis=0)
    num = np.einsum("nhm,hmd->nhd", qn, s) + n * v
    den = np.einsum("nhm,hm->nh", qn, z) + n
    o = num / den[..., None]
    gnum = g / den[..., None]
    gden = -(g * o).sum(-1) / den
    ds = np.einsum("nhm,nhd->hmd", qn, gnum)
    dz = np.einsum("nhm,nh->hm", qn, gden)
    dqn = np.einsum("nhd,hmd->nhm", gnum, s) + gden[..., None] * z[None]
    dv = n * gnum + np.einsum("nhm,hmd->nhd", kn, ds)
    dkn = np.einsum("nhd,hmd->nhm", v, ds) + dz[None]
    dq = (dqn - qn * (dqn * qn).sum()) / nq
    dk = (dkn - kn * (dkn * kn).sum()) / nk
    return dict(dq=dq, dk=dk, dv=dv, dS=ds, dz=dz)


def gcn_csr(edge_index, n):
    """Canonical CSR of the aggregation pattern: row index = edge target (`col`), column index = edge
    source (`row`), entries sorted by (target, source), duplicates kept (torch_sparse SparseTensor
    semantics).  Returns rowptr int64 [n+1], col int32 [nnz], dinv float32 [n]."""
    src = np.asarray(edge_index[0], dtype=np.int64)
    dst = np.asarray(edge_index[1], dtype=np.int64)
    order = np.lexsort((src, dst))
    deg = np.bincount(dst, minlength=n)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=rowptr[1:])
    # same operations, in fp32, as the reference: (1. / d).sqrt() with inf -> 0 (large/ours.py:29-32)
    with np.errstate(divide="ignore"):
        d32 = deg.astype(np.float32)
        dinv = np.where(deg > 0, np.sqrt(np.float32(1.0) / d32), np.float32(0.0)).astype(np.float32)
    return rowptr, src[order].astype(np.int32), dinv


def spmm_fp64(edge_index, n, x):
    """y = D^-1/2 A D^-1/2 x in float64 through scipy (duplicates summed == counted)."""
    src = np.asarray(edge_index[0], dtype=np.int64)
    dst = np.asarray(edge_index[1], dtype=np.int64)
    deg = np.bincount(dst, minlength=n).astype(np.float64)
    with np.errstate(divide="ignore"):
        dinv = np.where(deg > 0, 1.0 / np.sqrt(deg), 0.0)
    a = sp.coo_matrix((dinv[dst] * dinv[src], (dst, src)), shape=(n, n)).tocsr()
    return a @ np.asarray(x, dtype=np.float64)


# ---- graph preprocessing around the model (torch_geometric 1.7.2 semantics; integer, bit-exact) -------------------------------
def to_undirected(edge_index, n):
    """`torch_geometric.utils.to_undirected` as called at large/main.py:76 and medium/main.py:94: concatenate every edge
    with its reverse, then `torch_sparse.coalesce` = sort by row*n+col and drop duplicates.  int64 [2, nnz']."""
    ei = np.asarray(edge_index, dtype=np.int64)
    row = np.concatenate([ei[0], ei[1]])
    col = np.concatenate([ei[1], ei[0]])
    key = np.unique(row * np.int64(n) + col)          # sorted, deduplicated
    return np.stack([key // n, key % n]) if key.size else np.zeros((2, 0), dtype=np.int64)


def remove_self_loops(edge_index):
    """`remove_self_loops` (large/main.py:78, large/main-batch.py:97): mask = row != col, order preserved."""
    ei = np.asarray(edge_index, dtype=np.int64)
    return ei[:, ei[0] != ei[1]]


def add_self_loops(edge_index, n):
    """`add_self_loops(edge_index, num_nodes=n)` (large/main.py:79, large/main-batch.py:98): append (i, i) for i in 0..n-1."""
    ei = np.asarray(edge_index, dtype=np.int64)
    loops = np.arange(n, dtype=np.int64)
    return np.concatenate([ei, np.stack([loops, loops])], axis=1)


def eval_acc(y_true, y_pred_logits):
    """`eval_acc` of the reference (large/data_utils.py:210-220) for integer labels [m, k]: per label column, the fraction of
    rows whose argmax over the logits equals the label (every integer label counts as labelled); mean over columns."""
    y_true = np.asarray(y_true)
    y_pred = np.asarray(y_pred_logits).argmax(axis=-1)[:, None]
    accs = []
    for i in range(y_true.shape[1]):
        correct = y_true[:, i] == y_pred[:, i]
        accs.append(float(np.sum(correct)) / len(correct))
    return sum(accs) / len(accs)
