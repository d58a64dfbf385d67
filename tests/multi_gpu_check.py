"""torchrun script (not a pytest file): row-sharded execution on N GPUs over NCCL must reproduce the single-GPU result.
Run by tests/test_gpu_multi.py (pytest -m gpu on a box with >= 2 GPUs) and scripts/gpu_multi_*.sh:
    torchrun --nproc-per-node N tests/multi_gpu_check.py
Logits 2e-4 (fp32) / 2e-2 (bf16) of the largest logit; EVERY parameter gradient (weights and biases) within 3e-3 (fp32) /
8e-2 (bf16) of max(its own scale, 1e-3 of the model's largest gradient entry) - biases in front of a BatchNorm have an
analytically zero gradient, so their scale is pure rounding noise.  The sharded run sums every neighbourhood in a different order
(rotated column ids, fp32 partial sums between the SpMM phases); the weight gradients behind a BatchNorm amplify that
reordering noise to ~2e-3 of their scale - the same amplification the fp32 oracle shows against its fp64 run
(tests/test_gpu_model.py::test_model_matches_oracle_midsize) - measured 3.6e-4 (in-row waiting) / 2.5e-3 (phased) at world 2."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    from sgformer_b200 import large as L
    from sgformer_b200.dist import Comm
    from sgformer_b200.synth import make_graph
    ok = True
    for prec, tol in (("fp32", 2e-4), ("bf16", 2e-2)):
        torch.manual_seed(0)
        n, d, h, c = 30011, 48, 128, 7
        ei = make_graph(n, 200000, seed=3).to(dev)
        g = torch.Generator().manual_seed(1)
        x = torch.randn(n, d, generator=g).to(dev)
        y = torch.randint(0, c, (n,), generator=g).to(dev)
        kw = dict(gnn_num_layers=2, gnn_use_init=True, graph_weight=0.5, gnn_dropout=0.0, trans_dropout=0.0)
        ref = L.SGFormer(d, h, c, **kw).to(dev).set_precision(prec)
        for p in ref.parameters():
            dist.broadcast(p.data, 0)
        shard = L.SGFormer(d, h, c, **kw).to(dev).set_precision(prec)
        shard.load_state_dict(ref.state_dict())
        comm = Comm(dist.group.WORLD, n)
        shard.set_row_sharding(comm)
        r0, r1 = comm.rows
        ref.train(); shard.train()
        out_ref = ref(x, ei)
        (torch.nn.functional.nll_loss(torch.log_softmax(out_ref, 1), y, reduction="sum") / n).backward()
        out = shard(x[r0:r1].contiguous(), ei)
        (torch.nn.functional.nll_loss(torch.log_softmax(out, 1), y[r0:r1], reduction="sum") / n).backward()
        err = (out - out_ref[r0:r1]).abs().max().item() / out_ref.abs().max().item()
        gabs = max(q.grad.abs().max().item() for q in ref.parameters())
        gerr, worst = 0.0, ""
        for (k, p), (_, q) in zip(shard.named_parameters(), ref.named_parameters()):
            scale = max(q.grad.abs().max().item(), 1e-3 * gabs)
            e = (p.grad - q.grad).abs().max().item() / scale
            if e > gerr:
                gerr, worst = e, k
        good = err <= tol and gerr <= (3e-3 if prec == "fp32" else 8e-2)
        ok = ok and good
        if rank == 0 or not good:
            print(f"[{prec}] rank {rank} world={world}: logits rel err {err:.3e}, worst parameter-grad rel err {gerr:.3e} ({worst}) -> {'OK' if good else 'FAIL'}")
    # A schedule WITHOUT a collective between two uses of the same symmetric buffer (eval mode: no BatchNorm sums; no attention
    # layer: no Gram all-reduce; one GCN layer: buffer 0 every step) with the ranks deliberately skewed: a peer's push of step
    # t+1 must not land in the buffer (or raise a flag) before this rank has finished step t (Comm._fence_reuse).
    torch.manual_seed(0)
    kw0 = dict(trans_num_layers=0, gnn_num_layers=1, gnn_use_init=True, graph_weight=0.5, gnn_dropout=0.0, trans_dropout=0.0)
    ref = L.SGFormer(d, h, c, **kw0).to(dev)
    for p in ref.parameters():
        dist.broadcast(p.data, 0)
    shard = L.SGFormer(d, h, c, **kw0).to(dev)
    shard.load_state_dict(ref.state_dict())
    comm = Comm(dist.group.WORLD, n)
    shard.set_row_sharding(comm)
    ref.eval(); shard.eval()
    with torch.no_grad():
        xs = [x * (1.0 + 0.5 * t) for t in range(6)]
        locs = [v[r0:r1].contiguous() for v in xs]
        outs = []
        for t, v in enumerate(locs):
            if rank == t % world:
                torch.cuda._sleep(int(1e8))          # tens of ms of skew, no host sync anywhere in the loop
            outs.append(shard(v, ei))
        worst = 0.0
        for t, v in enumerate(xs):
            o = ref(v, ei)
            worst = max(worst, (outs[t] - o[r0:r1]).abs().max().item() / o.abs().max().item())
    good = worst <= 2e-4
    ok = ok and good
    if rank == 0 or not good:
        print(f"[no-collective eval loop] rank {rank}: logits rel err {worst:.3e} -> {'OK' if good else 'FAIL'}")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    passed = flag.item() == 1
    if rank == 0:
        print("multi_gpu_check: OK" if passed else "multi_gpu_check: FAILED", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    del shard, ref, comm          # symmetric buffers go before the process group they were rendezvoused on
    import gc
    gc.collect()
    dist.destroy_process_group()
    if not passed:
        sys.exit(1)


if __name__ == "__main__":
    main()
