"""sgformer_b200.optim.Adam (one fused launch per step, device step counter) against torch.optim.Adam with the reference's
two parameter groups (large/main.py:115-119), eager and replayed from a CUDA graph."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _params(seed, shapes):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g).to(DEV).requires_grad_(True) for s in shapes]


def test_adam_matches_torch_two_groups():
    from sgformer_b200.optim import Adam
    shapes = [(256, 100), (256,), (768, 256), (1,), (47, 256), (3, 5, 7)] + [(17,)] * 40      # > 32 tensors: two launches
    ours, ref = _params(0, shapes), _params(0, shapes)
    half = len(shapes) // 2
    kw = dict(lr=3e-3, betas=(0.85, 0.97), eps=1e-7)
    o1 = Adam([{"params": ours[:half], "weight_decay": 0.02}, {"params": ours[half:], "weight_decay": 0.0}], **kw)
    o2 = torch.optim.Adam([{"params": ref[:half], "weight_decay": 0.02}, {"params": ref[half:], "weight_decay": 0.0}], **kw)
    g = torch.Generator().manual_seed(1)
    for it in range(7):
        for a, b in zip(ours, ref):
            if it == 3 and a.shape == (1,):
                a.grad = b.grad = None          # parameters without a gradient are skipped
                continue
            gr = torch.randn(a.shape, generator=g).to(DEV) * (10.0 if it % 2 else 0.01)
            a.grad, b.grad = gr.clone(), gr.clone()
        o1.step()
        o2.step()
    for a, b in zip(ours, ref):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), f"{tuple(a.shape)}: {(a - b).abs().max().item():.3e}"
    st = o1.state_dict()["state"]
    assert set(st[0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and float(st[0]["step"]) == 7.0
    assert float(st[3]["step"]) == 6.0          # the parameter that had no gradient once


def test_adam_step_count_advances_under_cuda_graph():
    from sgformer_b200.optim import Adam
    p, q = _params(3, [(1000,)]), _params(3, [(1000,)])
    gfix = torch.randn(1000, generator=torch.Generator().manual_seed(4)).to(DEV)
    p[0].grad, q[0].grad = gfix.clone(), gfix.clone()
    o1, o2 = Adam(p, lr=1e-2), torch.optim.Adam(q, lr=1e-2)
    o1.step(); o2.step()                      # allocate state outside the capture
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        o1.step()
    torch.cuda.current_stream().wait_stream(side)
    o2.step()
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        o1.step()                              # captured, not executed
    for _ in range(5):
        cg.replay()
        o2.step()
    torch.cuda.synchronize()
    assert float(o1.state[p[0]]["step"]) == 7.0
    assert torch.allclose(p[0], q[0], rtol=5e-6, atol=1e-6), (p[0] - q[0]).abs().max().item()


def test_adam_rejects_cpu_parameters():
    from sgformer_b200.optim import Adam
    w = torch.zeros(4, requires_grad=True)
    w.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="CUDA"):
        Adam([w]).step()
