"""Property test (build container only): the oracle against the LIVE, unmodified reference modules on randomly drawn
configurations and graphs — isolated nodes, directed/asymmetric edges, duplicates, self loops, odd sizes, 1-4 heads
(SURVEY.md §4 'property' row).  The committed fixtures in tests/golden pin the same thing on the GPU box."""
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from _refload import build_reference_model, reference_available, run_reference
from oracle import sgformer_oracle as O

pytestmark = pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container)")


def _graph(n, e, seed, directed, isolated, dup, loops):
    g = torch.Generator().manual_seed(seed)
    hi = max(n - isolated, 1)
    ei = torch.stack([torch.randint(0, hi, (e,), generator=g), torch.randint(0, hi, (e,), generator=g)])
    if not directed:
        ei = torch.cat([ei, ei.flip(0)], 1)
    if dup:
        ei = torch.cat([ei, ei[:, :dup]], 1)
    if loops:
        ar = torch.arange(hi)
        ei = torch.cat([ei, torch.stack([ar, ar])], 1)
    return ei


@settings(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck))
@given(variant=st.sampled_from(["large", "100M", "medium"]), n=st.integers(5, 90), h=st.sampled_from([8, 16, 24]),
       heads=st.sampled_from([1, 2, 4]), seed=st.integers(0, 10 ** 6), directed=st.booleans(), isolated=st.integers(0, 3),
       dup=st.integers(0, 9), loops=st.booleans(), flags=st.lists(st.booleans(), min_size=9, max_size=9),
       aggregate=st.sampled_from(["add", "cat"]), layers=st.integers(1, 3), tlayers=st.integers(1, 2))
def test_oracle_equals_reference(variant, n, h, heads, seed, directed, isolated, dup, loops, flags, aggregate, layers, tlayers):
    d, c = 7, 4
    use_weight = flags[0] or heads > 1        # use_weight=False forces one head (medium/ours.py:84)
    if variant == "medium":
        cfg = O.make_config("medium", d, h, c, num_layers=tlayers, num_heads=heads, alpha=0.3, dropout=0.0, use_bn=flags[1],
                            use_residual=flags[2], use_weight=use_weight, gcn_num_layers=layers + 1, gcn_dropout=0.0,
                            gcn_use_bn=flags[3], graph_weight=0.7, aggregate=aggregate)
    else:
        kw = dict(trans_num_layers=tlayers, trans_num_heads=heads, trans_dropout=0.0, trans_use_bn=flags[1],
                  trans_use_residual=flags[2], trans_use_weight=use_weight, trans_use_act=flags[4], gnn_num_layers=layers,
                  gnn_dropout=0.0, gnn_use_weight=flags[5], gnn_use_init=flags[6], gnn_use_bn=flags[3], gnn_use_residual=flags[7],
                  gnn_use_act=flags[8], graph_weight=0.7, aggregate=aggregate)
        if variant == "100M":
            kw["alpha"] = 0.3
        cfg = O.make_config(variant, d, h, c, **kw)
    torch.manual_seed(seed)
    model, _ = build_reference_model(variant, cfg)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    ei = _graph(n, 3 * n, seed, directed, isolated, dup, loops)
    x = torch.randn(n, d, generator=torch.Generator().manual_seed(seed + 1))
    for training in (False, True):
        model.train(training)
        with torch.no_grad():
            ref = run_reference(variant, model, x, ei)
        model.load_state_dict(sd)   # undo BatchNorm running-stat updates
        out = O.sgformer_forward(cfg, sd, x, ei, training=training)
        err = (out - ref).abs().max().item()
        assert err <= 1e-4 * max(1.0, ref.abs().max().item()), f"training={training}: {err:.3e}"
