"""The Gram-form restatement of TransConvLayer (projections + full_attention_conv, medium/ours.py:14-34,76-95) that
sgformer_b200 runs for single-head layers, checked in fp64 against autograd of the reference formula.  The formulas live in
tests/kernel_emu.py (attn_gram_prepare_fwd / _bwd, ln_bwd_attn = the contracts of the CUDA kernels in csrc/attn_gram.cu);
this test pins them to the reference arithmetic, so that the GPU tests may compare the kernels with the emulation."""
import pytest
import torch

import kernel_emu as emu

DT = torch.float64


def _ref_layer(x, wq, bq, wk, bk, wv, bv):
    """Wq/Wk/Wv + full_attention_conv, H = 1 (medium/ours.py:76-95 with :14-34 inlined)."""
    n = x.shape[0]
    q, k, v = x @ wq.t() + bq, x @ wk.t() + bk, x @ wv.t() + bv
    qs, ks = q / torch.norm(q, p=2), k / torch.norm(k, p=2)
    kvs = ks.t() @ v
    num = qs @ kvs + n * v
    den = qs @ ks.sum(0) + n
    return num / den[:, None]


@pytest.mark.parametrize("n,h,use_weight", [(37, 12, True), (5, 8, True), (64, 16, False), (300, 24, True)])
def test_gram_form_matches_autograd(n, h, use_weight):
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, h, generator=g, dtype=DT).requires_grad_(True)
    ps = [(0.4 * torch.randn(h, h, generator=g, dtype=DT)).requires_grad_(True) if i % 2 == 0
          else torch.randn(h, generator=g, dtype=DT).requires_grad_(True) for i in range(4)]
    wq, bq, wk, bk = ps
    if use_weight:
        wv = (0.4 * torch.randn(h, h, generator=g, dtype=DT)).requires_grad_(True)
        bv = torch.randn(h, generator=g, dtype=DT).requires_grad_(True)
    else:
        wv, bv = torch.eye(h, dtype=DT).requires_grad_(True), torch.zeros(h, dtype=DT).requires_grad_(True)
    o = _ref_layer(x, wq, bq, wk, bk, wv, bv)
    gout = torch.randn(n, h, generator=g, dtype=DT)
    (o * gout).sum().backward()

    with torch.no_grad():
        xd = x.detach()
        G, s = xd.t() @ xd, xd.sum(0)
        st = emu.attn_gram_prepare_fwd(G, s, wq.detach(), bq.detach(), wk.detach(), bk.detach(), wv.detach(), bv.detach(), n)
        den = xd @ st.tail[0] + st.sc[emu.SC_DEN]
        o2 = (xd @ st.Bt.t() + st.bt) / den[:, None]
        assert (o2 - o).abs().max() < 1e-12 * max(1.0, o.abs().max().item())
        # row prologue (identity "LayerNorm": a = 1, no residual) + h x h backward
        gnum, gden, _, cs, pg, sg = emu.ln_bwd_attn(gout, o2, None, xd, 1.0, 0.0, None, None, None, False, False, 0.0, 0, 1.0, False,
                                                    None, None, den)
        P = xd.t() @ gnum
        dwq, dbq, dwk, dbk, dwv, dbv, bcat, a4 = emu.attn_gram_prepare_bwd(st, P, pg, cs, sg)
        dx = torch.cat([gnum, xd], 1) @ bcat.t() + torch.outer(gden, st.tail[0]) + a4
    for name, got, ref in [("dWq", dwq, wq.grad), ("dbq", dbq, bq.grad), ("dWk", dwk, wk.grad), ("dbk", dbk, bk.grad),
                           ("dWv", dwv, wv.grad), ("dbv", dbv, bv.grad), ("dx", dx, x.grad)]:
        scale = max(ref.abs().max().item(), 1e-30)
        assert (got - ref).abs().max().item() <= 1e-10 * scale + 1e-16, f"{name}: {(got - ref).abs().max().item():.3e} vs scale {scale:.3e}"
