#!/usr/bin/env python
"""K10 / K5 / K11 at a BASELINE shape on one GPU: device graph preprocessing (to_undirected -> remove_self_loops -> add_self_loops,
large/main.py:75-79), the CSR build, and the on-device evaluation, next to the numpy oracle of the same steps on the host.

    python scripts/bench_prep.py [--workload products] [--cpu]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="products")
    ap.add_argument("--cpu", action="store_true", help="also time the numpy oracle (tens of seconds at the products shape)")
    args = ap.parse_args()
    from sgformer_b200 import kernels as K
    from sgformer_b200.synth import SHAPES
    dev = torch.device("cuda:0")
    n, _, e, c, _, _, _ = SHAPES[args.workload]
    g = torch.Generator(device=dev).manual_seed(0)
    raw = torch.stack([torch.randint(0, n, (e,), generator=g, device=dev), torch.randint(0, n, (e,), generator=g, device=dev)])
    ms_u, und = timed(lambda: K.to_undirected(raw, n))
    ms_r, nsl = timed(lambda: K.remove_self_loops(und))
    ms_a, full = timed(lambda: K.add_self_loops(nsl, n))
    ms_c, csr = timed(lambda: K.csr_build(full, n))
    logits = torch.randn(n, c, device=dev)
    labels = torch.randint(0, c, (n, 1), device=dev)
    idx = torch.randperm(n, device=dev)[: n // 2]
    ms_e, acc = timed(lambda: K.eval_acc(logits, labels, idx, want_loss=True))
    print(f"{args.workload}: n={n} stored edges={e} -> undirected {und.shape[1]} -> +loops {full.shape[1]}")
    print(f"  GPU  to_undirected {ms_u:8.2f} ms | remove_self_loops {ms_r:7.2f} ms | add_self_loops {ms_a:7.2f} ms | csr_build {ms_c:7.2f} ms"
          f" | eval_acc+nll over {idx.numel()} rows {ms_e:6.3f} ms (incl. one scalar read-back each)")
    if args.cpu:
        from oracle import np_ref
        a = raw.cpu().numpy()
        t0 = time.perf_counter(); u = np_ref.to_undirected(a, n); t1 = time.perf_counter()
        r = np_ref.remove_self_loops(u); t2 = time.perf_counter()
        f = np_ref.add_self_loops(r, n); t3 = time.perf_counter()
        same = bool((torch.from_numpy(f) == full.cpu()).all())
        lg, lb, ix = logits.cpu(), labels.cpu(), idx.cpu()
        t4 = time.perf_counter(); acc_ref = np_ref.eval_acc(lb[ix].numpy(), lg[ix].numpy()); t5 = time.perf_counter()
        print(f"  host to_undirected {1e3 * (t1 - t0):8.1f} ms | remove_self_loops {1e3 * (t2 - t1):7.1f} ms | add_self_loops "
              f"{1e3 * (t3 - t2):7.1f} ms | eval_acc {1e3 * (t5 - t4):7.1f} ms   (numpy oracle, 1 thread)  bit-identical: {same}, "
              f"accuracy equal: {acc_ref == acc[0]}")


if __name__ == "__main__":
    main()
