"""Hand-scheduled forward / backward of the SGFormer encoder on the sm_100a kernels.

One schedule per branch — `trans_*` (TransConv: input MLP + linear-attention layers), `gconv_*` (GraphConv of
large/100M), `gcn_*` (PyG-GCN backbone of medium), `head_*` (branch mix + fc) — each a forward that records what the
backward needs in a `Tape`, and a backward that walks the tape and fills a gradient dict keyed by the reference's
parameter names.  There is no autograd inside: torch.autograd sees a single Function (functional.py) per call.

Reference semantics reproduced (paths into /root/reference): large/ours.py:25-42 (GraphConvLayer), :74-94 (GraphConv,
incl. the "residual always adds layer_[0]" quirk), :121-162 / medium/ours.py:14-46,74-100 (TransConvLayer +
full_attention_conv), :194-219 / medium/ours.py:133-160 / 100M/ours.py:247-272 (TransConv, the two residual rules),
medium/models.py:49-63 (GCN over PyG GCNConv), large/ours.py:265-276 (SGFormer.forward).
"""
from __future__ import annotations

import itertools
import os
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import kernels as K
from ._lib import EPI_ATTN_APPLY, EPI_ATTN_GRAM
from .dist import SINGLE, Comm
from .graph import Graph

Tensor = torch.Tensor


@dataclass
class Precision:
    """'bf16': bf16 activations, single-plane bf16 tensor-core operands (parity 1e-2).
    'fp32': fp32 activations, bf16x3 split operands — six tensor-core products per GEMM, fp32-accurate (parity 1e-4)."""
    name: str

    @property
    def act_dtype(self):
        return torch.bfloat16 if self.name == "bf16" else torch.float32

    @property
    def planes(self) -> int:
        return 1 if self.name == "bf16" else 3


BF16 = Precision("bf16")
FP32 = Precision("fp32")


def precision(name: str) -> Precision:
    if name in ("bf16", "bfloat16"):
        return BF16
    if name in ("fp32", "float32"):
        return FP32
    raise ValueError(f"unknown precision {name!r} (use 'bf16' or 'fp32')")


_seed_counter = itertools.count(1)


def next_seed() -> int:
    """Per-forward dropout seed: deterministic under torch.manual_seed, no device sync."""
    return ((torch.initial_seed() * 0x9E3779B1) ^ (next(_seed_counter) * 0x85EBCA6B)) & 0x7FFFFFFFFFFFFFFF


def check_width(h: int, prec: Precision, what: str):
    """The row kernels move rows in 16-byte chunks, at most 128 chunks per row (csrc/rowops.cu make_geom): fail with a clear
    message instead of SGF_ERR_ARG from the first LayerNorm/BatchNorm launch.  (The reference has no such limit.)"""
    vn = 8 if prec.name == "bf16" else 4
    if h % vn != 0 or h > 128 * vn:
        raise ValueError(f"sgformer_b200: {what} = {h} is not supported in precision '{prec.name}': it must be a multiple of {vn} "
                         f"and at most {128 * vn}" + (" (use set_precision('bf16') for wider layers)" if prec.name != "bf16" and h % 8 == 0 and h <= 1024 else ""))


class Tape(dict):
    """Saved tensors / scalars of one forward."""
    pass


def _w(P: Dict[str, Tensor], name: str, prec: Precision, transpose: bool = False) -> K.Operand:
    return K.pack_operand(P[name], transpose, prec.planes)


_xin_cache: "OrderedDict[tuple, tuple]" = OrderedDict()


def input_operand(x: Tensor, prec: Precision) -> K.Operand:
    """Raw node features fp32 [N, d_in] -> tensor-core operand (read by both branches' input Linear).  Full-batch
    training feeds the same feature tensor every step, so the packed operand is cached on the tensor's identity/version."""
    key = (x.data_ptr(), tuple(x.shape), x._version, str(x.dtype), prec.name, x.device.index)
    hit = _xin_cache.get(key)
    if hit is not None and hit[0] is x:
        _xin_cache.move_to_end(key)
        return hit[1]
    xs = x
    if xs.dtype != torch.float32:
        xs = xs.float()
    if xs.stride(-1) != 1:
        xs = xs.contiguous()
    op = K.pack_operand(xs.detach(), False, prec.planes)
    _xin_cache[key] = (x, op)
    while len(_xin_cache) > 2:
        _xin_cache.popitem(last=False)
    return op


# =================================================================================================
# linear attention core (full_attention_conv)
# =================================================================================================
def attention_forward(q: Tensor, k: Tensor, v: Tensor, heads: int, prec: Precision, tape: Optional[Tape],
                      comm: Comm = SINGLE, stats=None) -> Tensor:
    """q,k: [N, H*M], v: [N, H*D] activations (views allowed) -> o [N, H*D].  medium/ours.py:14-34.
    One Frobenius norm over all heads (medium/ours.py:16-17); N is the query count (the GLOBAL node count when the rows
    are sharded: the un-normalised partials {S', z', ||q||^2, ||k||^2} are all-reduced once, C1)."""
    n_loc = q.shape[0]
    n = comm.n_global if comm.active else n_loc
    m = q.shape[1] // heads
    d = v.shape[1] // heads
    dev = q.device
    if stats is not None:        # (sum of squares of q columns, column sums of k, sum of squares of k columns) from the
        sq_q, z_raw, sq_k = stats    # producing GEMM's epilogue
    else:
        _, sq_q = K.colstats(q, want_sum=False)
        z_raw, sq_k = K.colstats(k)
    s_list = []
    for hd in range(heads):
        kh, vh = k[:, hd * m:(hd + 1) * m], v[:, hd * d:(hd + 1) * d]
        s_raw = torch.empty((m, d), dtype=torch.float32, device=dev)
        K.gemm_tn(K.as_operand(kh, prec.planes, memo=True), K.as_operand(vh, prec.planes, memo=True), s_raw)
        s_list.append(s_raw)
    comm.allreduce_(sq_q, z_raw, sq_k, *s_list)
    o = K.alloc_act(n_loc, heads * d, q.dtype, dev)
    den = torch.empty((heads, n_loc), dtype=torch.float32, device=dev)
    scal = None
    for hd in range(heads):
        qh, vh = q[:, hd * m:(hd + 1) * m], v[:, hd * d:(hd + 1) * d]
        bmat, btail, scal = K.attn_prepare_fwd(s_list[hd], z_raw[hd * m:(hd + 1) * m], sq_q, sq_k, prec.planes)
        K.gemm_nt([K.as_operand(qh, prec.planes, memo=True)], [bmat], [(0, 0, 0, 0, m)], d, o[:, hd * d:(hd + 1) * d],
                  epi=EPI_ATTN_APPLY, aux=vh, tail=btail, nf=float(n), den_out=den[hd])
    if tape is not None:
        tape.update(q=q, k=k, v=v, o=o, den=den, s=s_list, z=z_raw, scal=scal, heads=heads, m=m, d=d, n=n)
    return o


def attention_backward(tape: Tape, g: Tensor, gscale: float, prec: Precision, dq: Tensor, dk: Tensor,
                       dv: Optional[Tensor], dv_accumulate: bool = False, comm: Comm = SINGLE):
    """g = dL/do [N, H*D] (times gscale).  Writes dq, dk [N, H*M] and dv [N, H*D] (+= if dv_accumulate).
    SURVEY.md Appendix A.1 in the raw-q/k form documented at sgf_attn_prepare_bwd; row-sharded: {dS', dz'} all-reduced (C2)."""
    q, k, v, o, den = tape["q"], tape["k"], tape["v"], tape["o"], tape["den"]
    heads, m, d, n = tape["heads"], tape["m"], tape["d"], tape["n"]
    dev = q.device
    scal_bwd = torch.zeros((heads, 8), dtype=torch.float32, device=dev)
    part = []
    for hd in range(heads):
        qh = q[:, hd * m:(hd + 1) * m]
        gnum, gden = K.attn_bwd_prep(g[:, hd * d:(hd + 1) * d], o[:, hd * d:(hd + 1) * d], den[hd], gscale)
        gnum_op = K.as_operand(gnum, prec.planes)
        ds_raw = torch.empty((m, d), dtype=torch.float32, device=dev)
        K.gemm_tn(K.as_operand(qh, prec.planes, memo=True), gnum_op, ds_raw)
        dz_raw, _ = K.colstats(qh, w=gden, want_sumsq=False)
        part.append((gnum, gden, gnum_op, ds_raw, dz_raw))
    comm.allreduce_(*[t for p_ in part for t in (p_[3], p_[4])])
    per_head = []
    for hd in range(heads):
        gnum, gden, gnum_op, ds_raw, dz_raw = part[hd]
        ops = K.attn_prepare_bwd(tape["s"][hd], tape["z"][hd * m:(hd + 1) * m], ds_raw, dz_raw, tape["scal"], prec.planes,
                                 scal_bwd[hd])
        per_head.append((gnum, gden, gnum_op, ops))
    if heads > 1:
        K.attn_combine_scal(scal_bwd, heads, tape["scal"])
    for hd in range(heads):
        gnum, gden, gnum_op, (b_dq, b_dv, b_dk, r1_col, dk_bias) = per_head[hd]
        qh, kh, vh = q[:, hd * m:(hd + 1) * m], k[:, hd * m:(hd + 1) * m], v[:, hd * d:(hd + 1) * d]
        sb = scal_bwd[hd]
        K.gemm_nt([gnum_op], [b_dq], [(0, 0, 0, 0, d)], m, dq[:, hd * m:(hd + 1) * m], alpha_dev=sb[0:1], aux=qh, beta=1.0,
                  beta_dev=sb[1:2], r1_row=gden, r1_col=r1_col)
        K.gemm_nt([K.as_operand(vh, prec.planes, memo=True)], [b_dk], [(0, 0, 0, 0, d)], m, dk[:, hd * m:(hd + 1) * m], alpha_dev=sb[0:1],
                  aux=kh, beta=1.0, beta_dev=sb[2:3], bias=dk_bias)
        if dv is not None:
            K.gemm_nt([K.as_operand(kh, prec.planes, memo=True)], [b_dv], [(0, 0, 0, 0, m)], d, dv[:, hd * d:(hd + 1) * d],
                      alpha_dev=sb[0:1], aux=gnum, beta=float(n), accumulate=dv_accumulate)


# =================================================================================================
# linear attention in Gram form (single head): projections + full_attention_conv without materialising q, k, v
# =================================================================================================
# With q = x Wq^T + bq, k = x Wk^T + bk, v = x Wv^T + bv every node-contracted quantity of full_attention_conv
# (medium/ours.py:16-31) is a function of the Gram matrix G = x^T x and the column sums s = x^T 1 of the layer input:
#   k^T v = Wk G Wv^T + (Wk s) bv^T + bk (Wv s)^T + N bk bv^T,   k^T 1 = Wk s + N bk,
#   ||k||^2 = <Wk G + bk s^T, Wk> + (k^T 1).bk   (same for q),
# so pass 1 is ONE node contraction x^T x (reads x once) and pass 2 ONE GEMM x . Bt^T with
#   Bt = (alpha/N) S'^T Wq + Wv,  out = (x Bt^T + bt) / (x ct + dt)      (alpha = 1/(||q|| ||k||); see sgf_attn_gram_prepare_fwd).
# The backward contracts P = x^T gnum', x^T gden' the same way: dWq, dWk, dWv come out of h x h algebra, dx is one
# two-segment GEMM [gnum' | x] . [Bt | A3] (sgf_attn_gram_prepare_bwd).  Forward traffic 3 N h b instead of 12 N h b,
# 4 N h^2 flops instead of 10 N h^2; identical mathematics (fp64 check: tests/test_gram_attention_math.py).
GRAM_ATTENTION = os.environ.get("SGF_GRAM_ATTENTION", "1") == "1"

_eye_cache: dict = {}


def _identity_v(h: int, dev):
    """use_weight=False: V is the layer input itself (medium/ours.py:84) = projection by the identity with zero bias."""
    key = (h, str(dev))
    if key not in _eye_cache:
        _eye_cache[key] = (torch.eye(h, dtype=torch.float32, device=dev), torch.zeros(h, dtype=torch.float32, device=dev))
    return _eye_cache[key]


def attention_gram_forward(P: Dict[str, Tensor], lp: str, x: Tensor, use_weight: bool, prec: Precision, tape: Optional[Tape],
                           comm: Comm = SINGLE) -> Tensor:
    """x: [N, h] layer input (activation) -> full_attention_conv(Wq x, Wk x, Wv x) [N, h], one head."""
    n_loc, h = x.shape
    n = comm.n_global if comm.active else n_loc
    dev = x.device
    xop = K.as_operand(x, prec.planes, memo=True)
    G, s = K.gram(xop, x)
    comm.allreduce_(G, s)                                     # C1: h*h + h floats
    wv, bv = (P[lp + "Wv.weight"], P[lp + "Wv.bias"]) if use_weight else _identity_v(h, dev)
    st = K.attn_gram_prepare_fwd(G, s, P[lp + "Wq.weight"], P[lp + "Wq.bias"], P[lp + "Wk.weight"], P[lp + "Wk.bias"], wv, bv, n)
    bop = K.pack_operand(st.Bt, False, prec.planes)
    btail = K.pack_operand(st.tail, False, prec.planes)
    d = st.Bt.shape[0]
    o = K.alloc_act(n_loc, d, x.dtype, dev)
    den = torch.empty(n_loc, dtype=torch.float32, device=dev)
    K.gemm_nt([xop], [bop], [(0, 0, 0, 0, h)], d, o, epi=EPI_ATTN_GRAM, bias=st.bt, tail=btail,
              nf_dev=st.sc[K.SC_DEN:K.SC_DEN + 1], den_out=den)
    if tape is not None:
        tape.update(st=st, den=den, xop=xop, n=n)
    return o


def attention_gram_backward(P: Dict[str, Tensor], lp: str, tape: Tape, x: Tensor, gnum: Tensor, gden: Tensor, cs: Tensor,
                            pg: Tensor, sg: Tensor, use_weight: bool, prec: Precision, dprev: Tensor, accumulate: bool,
                            grads: Dict[str, Tensor], comm: Comm = SINGLE):
    """gnum' = g/den~ [N,h], gden' = -(g.o)/den~ [N] and their column sums (sgf_ln_bwd_attn) -> parameter gradients and
    dprev (+)= d/dx of the attention (through q, k and v)."""
    st = tape["st"]
    h = x.shape[1]
    d = gnum.shape[1]
    dev = x.device
    gnum_op = K.as_operand(gnum, prec.planes)
    pmat = torch.empty((h, d), dtype=torch.float32, device=dev)
    K.gemm_tn(tape["xop"], gnum_op, pmat)
    comm.allreduce_(pmat, pg, cs, sg)                         # C2
    dwq, dbq, dwk, dbk, dwv, dbv, bcat, a4 = K.attn_gram_prepare_bwd(st, pmat, pg, cs, sg)
    grads[lp + "Wq.weight"], grads[lp + "Wq.bias"] = dwq, dbq
    grads[lp + "Wk.weight"], grads[lp + "Wk.bias"] = dwk, dbk
    names = [lp + "Wq.weight", lp + "Wq.bias", lp + "Wk.weight", lp + "Wk.bias"]
    if use_weight:
        grads[lp + "Wv.weight"], grads[lp + "Wv.bias"] = dwv, dbv
        names += [lp + "Wv.weight", lp + "Wv.bias"]
    _mark_global(grads, comm, *names)          # built from all-reduced contractions: already global sums
    bop = K.pack_operand(bcat, False, prec.planes)
    K.gemm_nt([gnum_op, tape["xop"]], [bop], [(0, 0, 0, 0, d), (1, 0, 0, d, h)], h, dprev, bias=a4, r1_row=gden,
              r1_col=st.tail[0], accumulate=accumulate)


# =================================================================================================
# TransConv branch
# =================================================================================================
def _res_coef(cfg: dict):
    if not cfg["trans_use_residual"]:
        return 1.0, 0.0, False
    if cfg["variant"] == "large":
        return 0.5, 0.5, True           # large/ours.py:211
    a = float(cfg["alpha"])
    return a, 1.0 - a, True             # medium/ours.py:152, 100M/ours.py:264


def _qkv_weight(P, pfx: str, use_weight: bool) -> (Tensor, Tensor):
    ws = [P[pfx + "Wq.weight"], P[pfx + "Wk.weight"]] + ([P[pfx + "Wv.weight"]] if use_weight else [])
    bs = [P[pfx + "Wq.bias"], P[pfx + "Wk.bias"]] + ([P[pfx + "Wv.bias"]] if use_weight else [])
    return torch.cat(ws, 0), torch.cat(bs, 0)


def trans_forward(P: Dict[str, Tensor], cfg: dict, xin: K.Operand, prec: Precision, training: bool, seed: int,
                  tape: Optional[Tape], pfx: str = "trans_conv.", comm: Comm = SINGLE) -> Tensor:
    h, H, d_in = cfg["hidden"], cfg["num_heads"], cfg["in_channels"]
    check_width(h, prec, "hidden_channels")
    n = xin.rows
    dev = xin.data.device
    p = float(cfg["trans_dropout"]) if training else 0.0
    use_ln = bool(cfg["trans_use_bn"])
    act = K.alloc_act(n, h, prec.act_dtype, dev)
    t0 = K.gemm_nt([xin], [_w(P, pfx + "fcs.0.weight", prec)], [(0, 0, 0, 0, d_in)], h, act, bias=P[pfx + "fcs.0.bias"])
    x, st = K.ln_fwd(t0, None, 1.0, 0.0, P.get(pfx + "bns.0.weight"), P.get(pfx + "bns.0.bias"), use_ln, True, p,
                     seed + 101, tape is not None)
    if tape is not None:
        tape.update(xin=xin, t0=t0, st0=st, layers=[], p=p, seed=seed, n=n)
    ca, cb, use_res = _res_coef(cfg)
    use_weight = bool(cfg["trans_use_weight"])
    if not use_weight and H != 1:
        raise ValueError("use_weight=False requires num_heads == 1 (medium/ours.py:84)")
    for i in range(cfg["trans_num_layers"]):
        lp = f"{pfx}convs.{i}."
        if GRAM_ATTENTION and H == 1:
            at = Tape() if tape is not None else None
            a = attention_gram_forward(P, lp, x, use_weight, prec, at, comm)
            y, st = K.ln_fwd(a, x if use_res else None, ca, cb, P.get(f"{pfx}bns.{i + 1}.weight"), P.get(f"{pfx}bns.{i + 1}.bias"),
                             use_ln, bool(cfg["trans_use_act"]), p, seed + 211 + i, tape is not None)
            if tape is not None:
                tape["layers"].append(dict(x_in=x, attn=at, a=a, st=st, gram=True))
            x = y
            continue
        wcat, bcat = _qkv_weight(P, lp, use_weight)
        nout = wcat.shape[0]
        qkv = torch.empty((n, K.ceil_to(nout, 8)), dtype=prec.act_dtype, device=dev)[:, :nout]
        csum = torch.zeros(nout, dtype=torch.float32, device=dev)
        csq = torch.zeros(nout, dtype=torch.float32, device=dev)
        K.gemm_nt([K.as_operand(x, prec.planes, memo=True)], [K.pack_operand(wcat, False, prec.planes)], [(0, 0, 0, 0, h)], nout, qkv,
                  bias=bcat, col_sum=csum, col_sumsq=csq)        # K^T 1, ||Q||^2, ||K||^2 fall out of the projection's epilogue
        q, k = qkv[:, :H * h], qkv[:, H * h:2 * H * h]
        v = qkv[:, 2 * H * h:] if use_weight else x
        at = Tape() if tape is not None else None
        o = attention_forward(q, k, v, H, prec, at, comm, stats=(csq[:H * h], csum[H * h:2 * H * h], csq[H * h:2 * H * h]))
        a = K.head_mean(o, H, h) if H > 1 else o
        y, st = K.ln_fwd(a, x if use_res else None, ca, cb, P.get(f"{pfx}bns.{i + 1}.weight"), P.get(f"{pfx}bns.{i + 1}.bias"),
                         use_ln, bool(cfg["trans_use_act"]), p, seed + 211 + i, tape is not None)
        if tape is not None:
            tape["layers"].append(dict(x_in=x, qkv=qkv, attn=at, a=a, st=st, nout=nout))
        x = y
    return x


def trans_attentions(P: Dict[str, Tensor], cfg: dict, xin: K.Operand, prec: Precision, with_act: bool,
                     pfx: str = "trans_conv.") -> List[Tensor]:
    """TransConv.get_attentions (large/ours.py:221-238, medium/ours.py:162-177, 100M/ours.py:274-289): per attention layer the
    [N, N] visualisation matrix  mean_h(q~_h k~_h^T) / mean_h(q~_h . sum_l k~_l + N)  of large/ours.py:152-155.  The N x N product is one
    tensor-core GEMM of the concatenated heads (sum over heads of per-head dot products) with 1/(H ||q|| ||k||) read from the
    device and the row normaliser as its row scale; the layer stack itself runs the un-fused attention (q, k materialised).
    Inference only (no dropout, no tape), O(N^2) memory like the reference: meant for small graphs."""
    h, H, d_in = cfg["hidden"], cfg["num_heads"], cfg["in_channels"]
    check_width(h, prec, "hidden_channels")
    n = xin.rows
    dev = xin.data.device
    use_ln = bool(cfg["trans_use_bn"])
    ca, cb, use_res = _res_coef(cfg)
    use_weight = bool(cfg["trans_use_weight"])
    t0 = K.gemm_nt([xin], [_w(P, pfx + "fcs.0.weight", prec)], [(0, 0, 0, 0, d_in)], h, K.alloc_act(n, h, prec.act_dtype, dev),
                   bias=P[pfx + "fcs.0.bias"])
    x, _ = K.ln_fwd(t0, None, 1.0, 0.0, P.get(pfx + "bns.0.weight"), P.get(pfx + "bns.0.bias"), use_ln, True, 0.0, 0, False)
    out = []
    for i in range(cfg["trans_num_layers"]):
        lp = f"{pfx}convs.{i}."
        wcat, bcat = _qkv_weight(P, lp, use_weight)
        nout = wcat.shape[0]
        qkv = torch.empty((n, K.ceil_to(nout, 8)), dtype=prec.act_dtype, device=dev)[:, :nout]
        K.gemm_nt([K.as_operand(x, prec.planes)], [K.pack_operand(wcat, False, prec.planes)], [(0, 0, 0, 0, h)], nout, qkv, bias=bcat)
        q, k = qkv[:, :H * h], qkv[:, H * h:2 * H * h]
        v = qkv[:, 2 * H * h:] if use_weight else x
        at = Tape()
        o = attention_forward(q, k, v, H, prec, at)
        inv_norm = at["den"].mean(dim=0).reciprocal_().contiguous()       # [N]: 1 / mean_h(den_h)  (an [N]-vector; not a hot path)
        att = K.alloc_act(n, n, torch.float32, dev)
        K.gemm_nt([K.as_operand(q, prec.planes)], [K.as_operand(k, prec.planes)], [(0, 0, 0, 0, H * h)], n, att, alpha=1.0 / H,
                  alpha_dev=at["scal"][2:3], row_scale=inv_norm)
        out.append(att)
        a = K.head_mean(o, H, h) if H > 1 else o
        x, _ = K.ln_fwd(a, x if use_res else None, ca, cb, P.get(f"{pfx}bns.{i + 1}.weight"), P.get(f"{pfx}bns.{i + 1}.bias"), use_ln,
                        with_act and bool(cfg["trans_use_act"]), 0.0, 0, False)
    return out


def trans_backward(P, cfg: dict, tape: Tape, dout: Tensor, gscale: float, prec: Precision, grads: Dict[str, Tensor],
                   pfx: str = "trans_conv.", want_dx: bool = False, comm: Comm = SINGLE) -> Optional[Tensor]:
    h, H, d_in = cfg["hidden"], cfg["num_heads"], cfg["in_channels"]
    n, p, seed = tape["n"], tape["p"], tape["seed"]
    dev = dout.device
    use_ln = bool(cfg["trans_use_bn"])
    ca, cb, use_res = _res_coef(cfg)
    use_weight = bool(cfg["trans_use_weight"])

    def zeros(k):
        return torch.zeros(k, dtype=torch.float32, device=dev)

    dcur, gs = dout, gscale
    for i in reversed(range(cfg["trans_num_layers"])):
        L = tape["layers"][i]
        lp = f"{pfx}convs.{i}."
        dg, db = (zeros(h), zeros(h)) if use_ln else (None, None)
        if L.get("gram"):
            x_in, at = L["x_in"], L["attn"]
            gnum, gden, dr, cs, pg, sg = K.ln_bwd_attn(dcur, L["a"], x_in if use_res else None, x_in, ca, cb,
                                                       P.get(f"{pfx}bns.{i + 1}.weight"), P.get(f"{pfx}bns.{i + 1}.bias"), L["st"],
                                                       use_ln, bool(cfg["trans_use_act"]), p, seed + 211 + i, gs, use_res, dg, db,
                                                       at["den"])
            if use_ln:
                grads[f"{pfx}bns.{i + 1}.weight"], grads[f"{pfx}bns.{i + 1}.bias"] = dg, db
            dprev = dr if dr is not None else K.new_like(x_in)
            attention_gram_backward(P, lp, at, x_in, gnum, gden, cs, pg, sg, use_weight, prec, dprev, dr is not None, grads, comm)
            dcur, gs = dprev, 1.0
            continue
        x_in, qkv, at, nout = L["x_in"], L["qkv"], L["attn"], L["nout"]
        da, dr = K.ln_bwd(dcur, L["a"], x_in if use_res else None, ca, cb, P.get(f"{pfx}bns.{i + 1}.weight"),
                          P.get(f"{pfx}bns.{i + 1}.bias"), L["st"], use_ln, bool(cfg["trans_use_act"]), p, seed + 211 + i, gs,
                          use_res, dg, db)
        if use_ln:
            grads[f"{pfx}bns.{i + 1}.weight"], grads[f"{pfx}bns.{i + 1}.bias"] = dg, db
        # head mean: every head receives da / H; da has pitch h, per-head slices of g are the same columns for all heads
        dqkv = torch.empty((n, K.ceil_to(nout, 8)), dtype=prec.act_dtype, device=dev)[:, :nout]
        dprev = dr
        if dprev is None:
            dprev = K.new_like(x_in)
            first_write = True
        else:
            first_write = False
        g_all = da if H == 1 else _tile_heads(da, H)
        if use_weight:
            attention_backward(at, g_all, 1.0 / H, prec, dqkv[:, :H * h], dqkv[:, H * h:2 * H * h], dqkv[:, 2 * H * h:],
                               comm=comm)
        else:
            # V is the layer input itself: its gradient goes straight into dprev
            attention_backward(at, g_all, 1.0 / H, prec, dqkv[:, :H * h], dqkv[:, H * h:2 * H * h], dprev,
                               dv_accumulate=not first_write, comm=comm)
            first_write = False
        wcat, _ = _qkv_weight(P, lp, use_weight)
        dqkv_op = K.as_operand(dqkv, prec.planes)
        K.gemm_nt([dqkv_op], [K.pack_operand(wcat, True, prec.planes)], [(0, 0, 0, 0, nout)], h, dprev,
                  accumulate=not first_write)
        dw = torch.empty((nout, h), dtype=torch.float32, device=dev)
        K.gemm_tn(dqkv_op, K.as_operand(x_in, prec.planes, memo=True), dw)
        dbias, _ = K.colstats(dqkv, want_sumsq=False)
        names = ["Wq", "Wk"] + (["Wv"] if use_weight else [])
        for j, nm in enumerate(names):
            grads[lp + nm + ".weight"] = dw[j * H * h:(j + 1) * H * h]
            grads[lp + nm + ".bias"] = dbias[j * H * h:(j + 1) * H * h]
        dcur, gs = dprev, 1.0
    dg, db = (zeros(h), zeros(h)) if use_ln else (None, None)
    dt0, _ = K.ln_bwd(dcur, tape["t0"], None, 1.0, 0.0, P.get(pfx + "bns.0.weight"), P.get(pfx + "bns.0.bias"), tape["st0"],
                      use_ln, True, p, seed + 101, gs, False, dg, db)
    if use_ln:
        grads[pfx + "bns.0.weight"], grads[pfx + "bns.0.bias"] = dg, db
    dt0_op = K.as_operand(dt0, prec.planes)
    dw0 = torch.empty((h, d_in), dtype=torch.float32, device=dev)
    K.gemm_tn(dt0_op, tape["xin"], dw0)
    grads[pfx + "fcs.0.weight"] = dw0
    grads[pfx + "fcs.0.bias"], _ = K.colstats(dt0, want_sumsq=False)
    if want_dx:
        dx = torch.empty((n, d_in), dtype=torch.float32, device=dev)
        K.gemm_nt([dt0_op], [_w(P, pfx + "fcs.0.weight", prec, transpose=True)], [(0, 0, 0, 0, h)], d_in, dx)
        return dx
    return None


def _tile_heads(da: Tensor, heads: int) -> Tensor:
    """[N, h] -> [N, H*h] with the same block repeated (gradient of the head mean, before the 1/H factor)."""
    n, h = da.shape
    out = K.alloc_act(n, heads * h, da.dtype, da.device)
    for hd in range(heads):
        K.axpby(da, None, 1.0, 0.0, out=out[:, hd * h:(hd + 1) * h])
    return out


# =================================================================================================
# GraphConv branch (large / 100M)
# =================================================================================================
def _stat_bufs(use_bn: bool, training: bool, h: int, dev):
    """Zeroed (sum, sumsq) buffers for a GEMM epilogue to fill when batch statistics are needed, else (None, None)."""
    if use_bn and training:
        return torch.zeros(h, dtype=torch.float32, device=dev), torch.zeros(h, dtype=torch.float32, device=dev)
    return None, None


def _bn_stats(z: Tensor, P, name: str, use_bn: bool, training: bool, zbias: Optional[Tensor] = None, comm: Comm = SINGLE,
              pre=None):
    if not use_bn:
        return None, None
    h = z.shape[1]
    if training:
        s, q = pre if (pre is not None and pre[0] is not None) else K.colstats(z)
        comm.allreduce_(s, q)                                              # C3: batch statistics span all shards
        rows = comm.n_global if comm.active else z.shape[0]
        mean, rstd = K.bn_finalize(s, q, rows, h, zbias, P.get(name + "running_mean"), P.get(name + "running_var"),
                                   z.device)
        nbt = P.get(name + "num_batches_tracked")
        if nbt is not None:
            nbt += 1
    else:
        mean, rstd = K.bn_finalize(None, None, z.shape[0], h, None, P[name + "running_mean"], P[name + "running_var"],
                                   z.device)
    return mean, rstd


def gconv_forward(P, cfg: dict, xin: K.Operand, graph: Graph, prec: Precision, training: bool, seed: int,
                  tape: Optional[Tape], mix: Optional[Tensor] = None, gw: float = 1.0, pfx: str = "graph_conv.",
                  comm: Comm = SINGLE) -> Tensor:
    """Returns GraphConv(x) — or, when `mix` is given, gw*GraphConv(x) + (1-gw)*mix (the SGFormer branch sum fused
    into the last layer's epilogue pass)."""
    h, d_in, nl = cfg["hidden"], cfg["in_channels"], cfg["gnn_num_layers"]
    check_width(h, prec, "hidden_channels")
    n = xin.rows
    dev = xin.data.device
    p = float(cfg["gnn_dropout"]) if training else 0.0
    use_bn, use_res, use_act = bool(cfg["gnn_use_bn"]), bool(cfg["gnn_use_residual"]), bool(cfg["gnn_use_act"])
    use_init, use_weight = bool(cfg["gnn_use_init"]), bool(cfg["gnn_use_weight"])
    dinv = graph.dinv
    st0 = _stat_bufs(use_bn, training, h, dev)
    z0 = K.gemm_nt([xin], [_w(P, pfx + "fcs.0.weight", prec)], [(0, 0, 0, 0, d_in)], h,
                   K.alloc_act(n, h, prec.act_dtype, dev), bias=P[pfx + "fcs.0.bias"], col_sum=st0[0], col_sumsq=st0[1])
    mean0, rstd0 = _bn_stats(z0, P, pfx + "bns.0.", use_bn, training, comm=comm, pre=st0)
    last_is_input = nl == 0
    # the pre-scaled SpMM operand is written where the halo exchange wants it (slot 0 of the step's symmetric buffer when pushed)
    x0, cur_s = K.bn_fwd(z0, None, mix if last_is_input else None, mean0, rstd0, P.get(pfx + "bns.0.weight"),
                         P.get(pfx + "bns.0.bias"), None, use_bn, True, p, seed + 307, gw, dinv, True, not last_is_input,
                         ys_out=None if last_is_input else comm.operand_out(n, h, prec.act_dtype, dev))
    if tape is not None:
        tape.update(xin=xin, z0=z0, mean0=mean0, rstd0=rstd0, x0=x0, layers=[], p=p, seed=seed, n=n, training=training,
                    mixed=mix is not None, gw=gw)
    out = x0
    for i in range(nl):
        last = i == nl - 1
        y = comm.spmm_gathered(K, graph, False, dinv, cur_s)    # C4: operand rows of every shard
        st = _stat_bufs(use_bn, training, h, dev)      # BatchNorm sums come out of the GEMM epilogue
        if use_init:
            w = _w(P, f"{pfx}convs.{i}.W.weight", prec)
            z = K.gemm_nt([K.as_operand(y, prec.planes, memo=True), K.as_operand(x0, prec.planes, memo=True)], [w],
                          [(0, 0, 0, 0, h), (1, 0, 0, h, h)], h, K.new_like(y), bias=P[f"{pfx}convs.{i}.W.bias"],
                          col_sum=st[0], col_sumsq=st[1])
        elif use_weight:
            w = _w(P, f"{pfx}convs.{i}.W.weight", prec)
            z = K.gemm_nt([K.as_operand(y, prec.planes, memo=True)], [w], [(0, 0, 0, 0, h)], h, K.new_like(y),
                          bias=P[f"{pfx}convs.{i}.W.bias"], col_sum=st[0], col_sumsq=st[1])
        else:
            z, st = y, (None, None)
        name = f"{pfx}bns.{i + 1}."
        mean, rstd = _bn_stats(z, P, name, use_bn, training, comm=comm, pre=st)
        yo, ys = K.bn_fwd(z, x0 if use_res else None, mix if last else None, mean, rstd, P.get(name + "weight"),
                          P.get(name + "bias"), None, use_bn, use_act, p, seed + 401 + i, gw, dinv, last, not last,
                          ys_out=None if last else comm.operand_out(n, h, prec.act_dtype, dev))
        if tape is not None:
            tape["layers"].append(dict(y=y, z=z, mean=mean, rstd=rstd))
        if last:
            out = yo
        else:
            cur_s = ys
    return out


def gconv_backward(P, cfg: dict, tape: Tape, graph: Graph, dout: Tensor, prec: Precision, grads: Dict[str, Tensor],
                   pfx: str = "graph_conv.", want_dx: bool = False, comm: Comm = SINGLE) -> Optional[Tensor]:
    """dout = gradient w.r.t. the tensor gconv_forward returned (the mixed tensor when `mix` was given: the factor gw
    is applied here; the caller routes (1-gw)*dout to the other branch)."""
    h, d_in, nl = cfg["hidden"], cfg["in_channels"], cfg["gnn_num_layers"]
    n, p, seed, training = tape["n"], tape["p"], tape["seed"], tape["training"]
    dev = dout.device
    use_bn, use_res, use_act = bool(cfg["gnn_use_bn"]), bool(cfg["gnn_use_residual"]), bool(cfg["gnn_use_act"])
    use_init, use_weight = bool(cfg["gnn_use_init"]), bool(cfg["gnn_use_weight"])
    dinv = graph.dinv
    rowptr_t, col_t = graph.transpose()
    x0 = tape["x0"]
    red = comm.allreduce_ if comm.active else None
    nstat = comm.n_global if comm.active else 0
    gs = tape["gw"] if tape["mixed"] else 1.0
    dx0 = None            # accumulated gradient of x0 (act dtype)
    dy_plain, dy_scaled = dout, None   # gradient entering the current layer's epilogue: plain, or pre-SpMM (needs *dinv)
    for i in reversed(range(nl)):
        L = tape["layers"][i]
        name = f"{pfx}bns.{i + 1}."
        if use_res and dx0 is None:
            dx0 = K.new_like(x0)
            res_acc = False
        else:
            res_acc = True
        dz, sums, colsum = K.bn_bwd(dy_plain, dy_scaled, dinv if dy_scaled is not None else None, L["z"], L["mean"],
                                    L["rstd"], P.get(name + "weight"), P.get(name + "bias"), None, use_bn, use_act, training,
                                    p, seed + 401 + i, gs, dres=dx0 if use_res else None, dres_accumulate=res_acc,
                                    want_dz_colsum=use_init or use_weight, reduce_fn=red, stat_rows=nstat)
        gs = 1.0
        if use_bn and training:
            grads[name + "bias"], grads[name + "weight"] = sums[:h], sums[h:]
            _mark_global(grads, comm, name + "bias", name + "weight")     # the BN sums were all-reduced between the phases
        elif use_bn:
            grads[name + "bias"], grads[name + "weight"] = _eval_bn_param_grads(dy_plain, dy_scaled, dinv, L, P, name, use_act)
        if use_init or use_weight:
            wname = f"{pfx}convs.{i}.W.weight"
            dz_op = K.as_operand(dz, prec.planes)
            kin = 2 * h if use_init else h
            dw = torch.empty((h, kin), dtype=torch.float32, device=dev)
            K.gemm_tn(dz_op, K.as_operand(L["y"], prec.planes, memo=True), dw[:, :h])
            if use_init:
                K.gemm_tn(dz_op, K.as_operand(x0, prec.planes, memo=True), dw[:, h:])
            grads[wname] = dw
            grads[f"{pfx}convs.{i}.W.bias"] = colsum
            wt = _w(P, wname, prec, transpose=True)   # [kin, h]
            dys = comm.operand_out(n, h, prec.act_dtype, dev)
            if dys is None or dys.stride(0) != dz.stride(0):
                dys = K.new_like(dz)
            K.gemm_nt([dz_op], [_slice_rows(wt, 0, h)], [(0, 0, 0, 0, h)], h, dys, row_scale=dinv)
            if use_init:
                if dx0 is None:
                    dx0 = K.new_like(x0)
                    K.gemm_nt([dz_op], [_slice_rows(wt, h, 2 * h)], [(0, 0, 0, 0, h)], h, dx0)
                else:
                    K.gemm_nt([dz_op], [_slice_rows(wt, h, 2 * h)], [(0, 0, 0, 0, h)], h, dx0, accumulate=True)
        else:
            dys = K.axpby(dz, None, 1.0, 0.0, row_scale=dinv, out=comm.operand_out(n, h, prec.act_dtype, dev))
        # = A^T (dinv . dy): gradient w.r.t. the pre-scaled SpMM input (C4: gradient rows of every shard)
        dy_scaled = comm.spmm_gathered(K, graph, True, None, dys)
        dy_plain = None
    # input layer epilogue: gradient of x0 = accumulated dx0 (+ dinv * dy_scaled from layer 0's SpMM)
    if nl == 0:
        g_plain, g_scaled = dout, None
    else:
        g_plain, g_scaled = dx0, dy_scaled
    dz0, sums, colsum = K.bn_bwd(g_plain, g_scaled, dinv if g_scaled is not None else None, tape["z0"], tape["mean0"],
                                 tape["rstd0"], P.get(pfx + "bns.0.weight"), P.get(pfx + "bns.0.bias"), None, use_bn, True,
                                 training, p, seed + 307, gs, want_dz_colsum=True, reduce_fn=red, stat_rows=nstat)
    if use_bn and training:
        grads[pfx + "bns.0.bias"], grads[pfx + "bns.0.weight"] = sums[:h], sums[h:]
        _mark_global(grads, comm, pfx + "bns.0.bias", pfx + "bns.0.weight")
    elif use_bn:
        grads[pfx + "bns.0.bias"], grads[pfx + "bns.0.weight"] = _eval_bn_param_grads(
            g_plain, g_scaled, dinv, dict(z=tape["z0"], mean=tape["mean0"], rstd=tape["rstd0"]), P, pfx + "bns.0.", True)
    dz0_op = K.as_operand(dz0, prec.planes)
    dw0 = torch.empty((h, d_in), dtype=torch.float32, device=dev)
    K.gemm_tn(dz0_op, tape["xin"], dw0)
    grads[pfx + "fcs.0.weight"] = dw0
    grads[pfx + "fcs.0.bias"] = colsum
    if want_dx:
        dx = torch.empty((n, d_in), dtype=torch.float32, device=dev)
        K.gemm_nt([dz0_op], [_w(P, pfx + "fcs.0.weight", prec, transpose=True)], [(0, 0, 0, 0, h)], d_in, dx)
        return dx
    return None


def _mark_global(grads: dict, comm: Comm, *names: str):
    """Gradients that are already sums over all shards (excluded from the C5 all-reduce)."""
    if comm.active:
        grads.setdefault("__global__", set()).update(names)


def _eval_bn_param_grads(dy, dy2, dinv, L, P, name, use_relu):
    """BatchNorm affine gradients in eval mode (running statistics; rare: eval-mode backward)."""
    h = L["z"].shape[1]
    sums = K.bn_bwd_sums(dy, dy2, dinv if dy2 is not None else None, L["z"], L["mean"], L["rstd"], P[name + "weight"],
                         P[name + "bias"], None, True, use_relu, 0.0, 0, 1.0)
    return sums[:h], sums[h:]


def _slice_rows(op: K.Operand, r0: int, r1: int) -> K.Operand:
    return K.Operand(op.data[r0:r1], r1 - r0, op.k, op.kp, op.planes)


# =================================================================================================
# GCN backbone (medium): PyG GCNConv stack
# =================================================================================================
def gcn_forward(P, cfg: dict, xin: K.Operand, graph: Graph, prec: Precision, training: bool, seed: int,
                tape: Optional[Tape], mix: Optional[Tensor] = None, gw: float = 1.0, pfx: str = "gnn.",
                comm: Comm = SINGLE) -> Tensor:
    """models.GCN.forward (medium/models.py:49-63).  `graph` is built with PyG self-loop semantics (gcn_norm)."""
    nl = cfg["gcn_num_layers"]
    n = xin.rows
    dev = xin.data.device
    p = float(cfg["gcn_dropout"]) if training else 0.0
    use_bn = bool(cfg["gcn_use_bn"])
    dinv = graph.dinv
    cur_op, cur_k = xin, cfg["in_channels"]
    layers = []
    out = None
    for i in range(nl):
        last = i == nl - 1
        wname = f"{pfx}convs.{i}.lin.weight"
        hout = P[wname].shape[0]
        check_width(hout, prec, f"GCN layer {i} out_channels")
        t = K.gemm_nt([cur_op], [_w(P, wname, prec)], [(0, 0, 0, 0, cur_k)], hout, K.alloc_act(n, hout, prec.act_dtype, dev),
                      row_scale=dinv)
        s = comm.spmm_gathered(K, graph, False, dinv, t)
        zb = P.get(f"{pfx}convs.{i}.bias")
        if last:
            out, _ = K.bn_fwd(s, None, mix, None, None, None, None, zb, False, False, 0.0, 0, gw, None, True, False)
            layers.append(dict(s=s, cur_op=cur_op, cur_k=cur_k, hout=hout))
        else:
            name = f"{pfx}bns.{i}."
            mean, rstd = _bn_stats(s, P, name, use_bn, training, zbias=zb, comm=comm) if use_bn else (None, None)
            y, _ = K.bn_fwd(s, None, None, mean, rstd, P.get(name + "weight"), P.get(name + "bias"), zb, use_bn, True, p,
                            seed + 503 + i, 1.0, None, True, False)
            layers.append(dict(s=s, cur_op=cur_op, cur_k=cur_k, hout=hout, mean=mean, rstd=rstd))
            cur_op, cur_k = K.as_operand(y, prec.planes), hout
    if tape is not None:
        tape.update(layers=layers, p=p, seed=seed, n=n, training=training, mixed=mix is not None, gw=gw)
    return out


def gcn_backward(P, cfg: dict, tape: Tape, graph: Graph, dout: Tensor, prec: Precision, grads: Dict[str, Tensor],
                 pfx: str = "gnn.", want_dx: bool = False, comm: Comm = SINGLE) -> Optional[Tensor]:
    nl = cfg["gcn_num_layers"]
    n, p, seed, training = tape["n"], tape["p"], tape["seed"], tape["training"]
    dev = dout.device
    use_bn = bool(cfg["gcn_use_bn"])
    dinv = graph.dinv
    rowptr_t, col_t = graph.transpose()
    red = comm.allreduce_ if comm.active else None
    nstat = comm.n_global if comm.active else 0
    gs = tape["gw"] if tape["mixed"] else 1.0
    dcur = dout
    dx = None
    for i in reversed(range(nl)):
        L = tape["layers"][i]
        last = i == nl - 1
        zb = P.get(f"{pfx}convs.{i}.bias")
        hout = L["hout"]
        if last:
            dzs, _, colsum = K.bn_bwd(dcur, None, None, L["s"], None, None, None, None, zb, False, False, training, 0.0, 0, gs,
                                      want_dz_colsum=zb is not None, out_row_scale=dinv)
        else:
            name = f"{pfx}bns.{i}."
            dzs, sums, colsum = K.bn_bwd(dcur, None, None, L["s"], L.get("mean"), L.get("rstd"), P.get(name + "weight"),
                                         P.get(name + "bias"), zb, use_bn, True, training, p, seed + 503 + i, gs,
                                         want_dz_colsum=zb is not None, out_row_scale=dinv, reduce_fn=red, stat_rows=nstat)
            if use_bn and training:
                grads[name + "bias"], grads[name + "weight"] = sums[:hout], sums[hout:]
                _mark_global(grads, comm, name + "bias", name + "weight")
        gs = 1.0
        if zb is not None:
            grads[f"{pfx}convs.{i}.bias"] = colsum
        u = comm.spmm_gathered(K, graph, True, dinv, dzs)        # = Â^T dz = gradient of (x W^T)
        u_op = K.as_operand(u, prec.planes)
        wname = f"{pfx}convs.{i}.lin.weight"
        dw = torch.empty((hout, L["cur_k"]), dtype=torch.float32, device=dev)
        K.gemm_tn(u_op, L["cur_op"], dw)
        grads[wname] = dw
        if i > 0:
            dcur = K.alloc_act(n, L["cur_k"], prec.act_dtype, dev)
            K.gemm_nt([u_op], [_w(P, wname, prec, transpose=True)], [(0, 0, 0, 0, hout)], L["cur_k"], dcur)
        elif want_dx:
            dx = torch.empty((n, L["cur_k"]), dtype=torch.float32, device=dev)
            K.gemm_nt([u_op], [_w(P, wname, prec, transpose=True)], [(0, 0, 0, 0, hout)], L["cur_k"], dx)
    return dx


# =================================================================================================
# head: fc over the mixed / concatenated branches
# =================================================================================================
def head_forward(P, cfg: dict, feats: List[Tensor], prec: Precision, tape: Optional[Tape]) -> Tensor:
    """feats = [m] ('add', branches already mixed) or [x1, x2] ('cat').  large/ours.py:269-275.  Logits are fp32."""
    h, c = cfg["hidden"], cfg["out_channels"]
    n = feats[0].shape[0]
    w = _w(P, "fc.weight", prec)
    ops = [K.as_operand(f, prec.planes) for f in feats]
    pairs = [(j, 0, 0, j * h, h) for j in range(len(feats))]
    # pitch padded to a 16-byte multiple (c = 47 -> 48 floats): the GEMM epilogue can then use its TMA-store path
    out = K.alloc_act(n, c, torch.float32, feats[0].device)
    K.gemm_nt(ops, [w], pairs, c, out, bias=P["fc.bias"])
    if tape is not None:
        tape.update(ops=ops, nfeat=len(feats))
    return out


def head_backward(P, cfg: dict, tape: Tape, dlogits: Tensor, prec: Precision, grads: Dict[str, Tensor]) -> List[Tensor]:
    h, c = cfg["hidden"], cfg["out_channels"]
    n = dlogits.shape[0]
    dev = dlogits.device
    dlogits = dlogits.contiguous().float()
    db = torch.zeros(c, dtype=torch.float32, device=dev)
    dl_op = K.pack_operand(dlogits, False, prec.planes, colsum=db)
    nf = tape["nfeat"]
    dw = torch.empty((c, nf * h), dtype=torch.float32, device=dev)
    wt = _w(P, "fc.weight", prec, transpose=True)   # [nf*h, c]
    outs = []
    for j in range(nf):
        K.gemm_tn(dl_op, tape["ops"][j], dw[:, j * h:(j + 1) * h])
        dj = K.alloc_act(n, h, prec.act_dtype, dev)
        K.gemm_nt([dl_op], [_slice_rows(wt, j * h, (j + 1) * h)], [(0, 0, 0, 0, c)], h, dj)
        outs.append(dj)
    grads["fc.weight"], grads["fc.bias"] = dw, db
    return outs
