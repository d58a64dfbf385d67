"""The reference's mini-batch training loop, as its driver runs it, on top of the drop-in module on a real GPU.

Mirrors the semantics of large/main-batch.py:118-158 and large/eval.py:35-65 (`evaluate_large(device="cpu")`) without the
dataset / logger plumbing: `model.reset_parameters()`, a two-group `torch.optim.Adam` built from `model.params1 / params2`,
per epoch `model.to(device); model.train()`, a CPU `randperm` cut into batches, per batch `x[idx].to(device)`, the CPU
`subgraph(idx, edge_index, num_nodes=n, relabel_nodes=True)` (torch_geometric restatement of tests/ref_shims), `log_softmax` +
`NLLLoss` on the batch's training rows, and between epochs the evaluation that moves the model to the CPU and calls it with CPU
tensors.  The graph is a planted-partition graph whose classes are learnable, so the loop must actually train."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _planted(n, c, d, deg, seed):
    g = torch.Generator().manual_seed(seed)
    y = torch.randint(0, c, (n,), generator=g)
    centers = torch.randn(c, d, generator=g)
    x = centers[y] + 1.5 * torch.randn(n, d, generator=g)
    src = torch.randint(0, n, (n * deg,), generator=g)
    same = torch.rand(n * deg, generator=g) < 0.8                      # 80 % of the edges stay inside the class
    order = torch.argsort(y)
    starts = torch.searchsorted(y[order], torch.arange(c))
    counts = torch.bincount(y, minlength=c)
    pick = starts[y[src]] + (torch.rand(n * deg, generator=g) * counts[y[src]]).long().clamp_max(n - 1)
    dst = torch.where(same, order[pick.clamp_max(n - 1)], torch.randint(0, n, (n * deg,), generator=g))
    ei = torch.stack([torch.cat([src, dst]), torch.cat([dst, src])])
    loops = torch.arange(n)
    return x, torch.cat([ei, torch.stack([loops, loops])], 1), y.unsqueeze(1)


def test_reference_minibatch_loop_trains_on_the_dropin():
    sys.path.insert(0, os.path.join(HERE, "ref_shims"))
    try:
        from torch_geometric.utils import subgraph          # CPU restatement of PyG 1.7.2 (tests/ref_shims)
    finally:
        sys.path.pop(0)
    from sgformer_b200.large import SGFormer                # what `from ours import *` resolves to under sgformer_b200.launch

    device = torch.device("cuda:0")
    torch.manual_seed(0)
    n, c, d, batch_size, epochs = 20000, 5, 32, 5000, 5
    x, edge_index, true_label = _planted(n, c, d, 6, 1)
    perm = torch.randperm(n)
    split_idx = {"train": perm[: n // 2], "valid": perm[n // 2: 3 * n // 4], "test": perm[3 * n // 4:]}
    train_mask = torch.zeros(n, dtype=torch.bool)
    train_mask[split_idx["train"]] = True
    model = SGFormer(d, 64, c, trans_num_layers=1, trans_dropout=0.2, gnn_num_layers=2, gnn_dropout=0.2, gnn_use_init=True,
                     graph_weight=0.5).to(device)
    model.reset_parameters()
    optimizer = torch.optim.Adam([{"params": model.params1, "weight_decay": 1e-3}, {"params": model.params2, "weight_decay": 5e-4}],
                                 lr=0.01)
    criterion = torch.nn.NLLLoss()
    num_batch = n // batch_size + (n % batch_size > 0)
    epoch_loss, accs = [], []
    for epoch in range(epochs):
        model.to(device)
        model.train()
        idx = torch.randperm(n)
        losses = []
        for i in range(num_batch):
            idx_i = idx[i * batch_size:(i + 1) * batch_size]
            train_mask_i = train_mask[idx_i]
            x_i = x[idx_i].to(device)
            edge_index_i, _ = subgraph(idx_i, edge_index, num_nodes=n, relabel_nodes=True)
            edge_index_i = edge_index_i.to(device)
            y_i = true_label[idx_i].to(device)
            optimizer.zero_grad()
            out_i = F.log_softmax(model(x_i, edge_index_i), dim=1)
            loss = criterion(out_i[train_mask_i], y_i.squeeze(1)[train_mask_i])
            loss.backward()
            optimizer.step()
            losses.append(float(loss))
        epoch_loss.append(sum(losses) / len(losses))
        # evaluate_large(..., device="cpu"): the model and the data go to the CPU, the forward is called with CPU tensors
        with torch.no_grad():
            model.eval()
            model.to(torch.device("cpu"))
            out = model(x, edge_index)
            assert out.device.type == "cpu" and out.shape == (n, c)
            pred = out.argmax(dim=1)
            accs.append({k: float((pred[v] == true_label.squeeze(1)[v]).float().mean()) for k, v in split_idx.items()})
    assert all(l == l for l in epoch_loss), epoch_loss
    assert epoch_loss[-1] < 0.6 * epoch_loss[0], f"the loop did not train: epoch losses {epoch_loss}"
    assert accs[-1]["test"] > 0.85 and accs[-1]["valid"] > 0.85, accs
    # the CPU-called evaluation is the same computation as a CUDA call
    model.to(device)
    with torch.no_grad():
        out_dev = model(x.to(device), edge_index.to(device))
    assert torch.allclose(out_dev.cpu(), out, rtol=1e-4, atol=1e-4)
