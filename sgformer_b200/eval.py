"""`evaluate()` of the reference drivers with the metric computed on the device.

Reference: large/eval.py:6-33 runs the model, then for each of the three splits gathers `out[split]`, takes the argmax, copies both
label and prediction arrays to the host and loops in numpy (`eval_acc`, large/data_utils.py:210-220); the validation loss is
`NLLLoss(log_softmax(out)[valid], label[valid])`.  Here one kernel launch per split (`sgf_eval_acc`, K11) reads the logits in
place through the split's index list and returns the hit count (and, for the validation split, the summed NLL); only those
scalars cross PCIe.  Same signature and return tuple as the reference, so `large/main.py:144` can call it unchanged:

    from sgformer_b200.eval import evaluate

Metrics other than `eval_acc` (rocauc / f1: host-side sklearn in the reference) and the multi-label datasets are delegated to the
caller's own `eval_func` / `criterion` exactly as the reference does.
"""
from __future__ import annotations

import torch

from . import kernels as K

_BCE_DATASETS = ('yelp-chi', 'deezer-europe', 'twitch-e', 'fb100', 'ogbn-proteins')      # large/eval.py:21


@torch.no_grad()
def evaluate(model, dataset, split_idx, eval_func, criterion, args, result=None):
    if result is not None:
        out = result
    else:
        model.eval()
        out = model(dataset.graph['node_feat'], dataset.graph['edge_index'])
    label = dataset.label
    on_device = (getattr(eval_func, "__name__", "") == "eval_acc" and label.dim() == 2 and label.shape[1] == 1
                 and label.dtype == torch.int64 and getattr(args, "dataset", None) not in _BCE_DATASETS)
    if not on_device:
        # other metrics (rocauc, f1: sklearn on the host in the reference) and the multi-label datasets: the caller's own
        # eval_func / criterion, exactly as large/eval.py:13-31 - these are outside the accelerated path, not a fallback of K11
        train_acc = eval_func(label[split_idx['train']], out[split_idx['train']])
        valid_acc = eval_func(label[split_idx['valid']], out[split_idx['valid']])
        test_acc = eval_func(label[split_idx['test']], out[split_idx['test']])
        if getattr(args, "dataset", None) in _BCE_DATASETS:
            true_label = torch.nn.functional.one_hot(label, label.max() + 1).squeeze(1) if label.shape[1] == 1 else label
            valid_loss = criterion(out[split_idx['valid']], true_label.squeeze(1)[split_idx['valid']].to(torch.float))
        else:
            out = torch.log_softmax(out, dim=1)
            valid_loss = criterion(out[split_idx['valid']], label.squeeze(1)[split_idx['valid']])
        return train_acc, valid_acc, test_acc, valid_loss, out
    logits = out.float()
    if not logits.is_cuda:              # logits handed in from the host (e.g. `result=` of evaluate_large): K11 has no CPU path
        logits = logits.cuda()
    lab = label.to(logits.device)
    train_acc, _ = K.eval_acc(logits, lab, split_idx['train'])
    valid_acc, valid_loss = K.eval_acc(logits, lab, split_idx['valid'], want_loss=True)
    test_acc, _ = K.eval_acc(logits, lab, split_idx['test'])
    return train_acc, valid_acc, test_acc, torch.tensor(valid_loss, dtype=torch.float32), torch.log_softmax(out, dim=1)
