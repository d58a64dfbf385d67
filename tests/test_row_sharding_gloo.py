"""Row-sharded schedule (C1-C5) on world_size 2 and 3 with the gloo backend on CPU.

Each rank runs the fused forward/backward on its row block with the kernels replaced by their torch-CPU contract
(tests/kernel_emu.py); the concatenated logits and the all-reduced parameter gradients must equal the single-process
run.  Integer shard bookkeeping (row partition, shard CSR) is checked bit-exactly against slicing the global CSR."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cfg_from_oracle(c):
    from sgformer_b200.config import make_config
    keys = make_config("large", 1, 1, 1).keys()
    kw = {k: v for k, v in c.items() if k in keys and k not in ("variant", "in_channels", "hidden", "out_channels")}
    return make_config(c["variant"], c["in_channels"], c["hidden"], c["out_channels"], **kw)


def _worker(rank, world, port, fixture, outdir, c4_mode="allgather"):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    import kernel_emu
    from sgformer_b200 import engine as E
    from sgformer_b200 import functional as Fn
    from sgformer_b200 import dist as D
    from sgformer_b200.dist import Comm
    E.K = kernel_emu
    Fn.K = kernel_emu
    D.C4_CHUNKS, D.C4_MIN_CHUNK_BYTES = 2, 8     # exercise the column-chunked all-gather / SpMM pipeline on the tiny fixtures
    fx = torch.load(fixture, weights_only=False)
    cfg = _cfg_from_oracle(fx["cfg"])
    sd = fx["state_dict"]
    names = tuple(sd.keys())
    n = fx["x"].shape[0]
    comm = Comm(dist.group.WORLD, n, c4_mode=c4_mode)
    r0, r1 = comm.rows
    graph = kernel_emu.EmuGraph(fx["edge_index"], n, 1 if cfg["variant"] == "medium" else 0, rows=(r0, r1), col_rot=comm.col_rot)
    params = [sd[k].clone().requires_grad_(True) if (sd[k].is_floating_point() and "running" not in k) else sd[k].clone()
              for k in names]
    x = fx["x"][r0:r1].clone().requires_grad_(True)
    out = Fn.SGFormerFn.apply(x, graph, cfg, E.FP32, True, comm, names, *params)
    (out * fx["loss_weight"][r0:r1]).sum().backward()
    torch.save(dict(rows=(r0, r1), out=out.detach(), grad_x=x.grad,
                    grads={k: p.grad for k, p in zip(names, params) if getattr(p, "grad", None) is not None},
                    buffers={k: p for k, p in zip(names, params) if "running" in k}), os.path.join(outdir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def _close(a, b, rtol, atol, what):
    a, b = a.double(), b.double()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    assert err <= atol + rtol * ref, f"{what}: max err {err:.3e} (ref max {ref:.3e})"


@pytest.mark.parametrize("world,c4_mode", [(2, "allgather"), (3, "allgather"), (2, "rotated"), (3, "rotated")])
@pytest.mark.parametrize("name", ["large_add_init", "large_cat_heads2", "100M_alpha", "medium_gcn"])
def test_row_sharded_matches_single_process(tmp_path, world, c4_mode, name):
    """c4_mode 'rotated' = the slot layout of the pushed halo exchange (rotated column ids in the CSR shard, operand blocks in
    arrival order) filled by an all-gather; the copy-engine push itself needs GPUs (tests/test_gpu_multi.py)."""
    fixture = os.path.join(GOLD, f"model_{name}.pt")
    mp.spawn(_worker, args=(world, _free_port(), fixture, str(tmp_path), c4_mode), nprocs=world, join=True)
    fx = torch.load(fixture, weights_only=False)
    parts = [torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"), weights_only=False) for r in range(world)]
    out = torch.cat([p["out"] for p in parts])
    _close(out, fx["out_train"], 5e-5, 5e-6, "sharded train logits")
    _close(torch.cat([p["grad_x"] for p in parts]), fx["grad_x"], 1e-3, 5e-6, "sharded grad x")
    for k, g in fx["grads"].items():
        for r, p in enumerate(parts):
            _close(p["grads"][k], g, 1e-3, 5e-5, f"rank {r} grad {k} (all-reduced)")
    for k, v in fx["buffers_after_train"].items():
        if "running" in k:
            _close(parts[0]["buffers"][k].float(), v.float(), 1e-4, 1e-5, f"buffer {k}")


def test_partition_and_shard_csr_are_exact():
    sys.path.insert(0, HERE)
    import kernel_emu
    from sgformer_b200.dist import partition
    n, world = 103, 4
    blocks = [partition(n, world, r) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n and all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
    g = torch.Generator().manual_seed(0)
    ei = torch.randint(0, n, (2, 900), generator=g)
    full = kernel_emu.csr_build(ei, n)
    block = blocks[0][1] - blocks[0][0]
    for (r0, r1) in blocks:
        rp, cl, dv = kernel_emu.csr_build(ei, n, rows=(r0, r1))
        assert torch.equal(rp, full[0][r0:r1 + 1] - full[0][r0])
        assert torch.equal(cl, full[1][full[0][r0]:full[0][r1]])
        assert torch.equal(dv, full[2][r0:r1])
        # rotated storage: same multiset per row, ids shifted so that the shard's own block comes first, rows sorted
        rpr, clr, dvr = kernel_emu.csr_build(ei, n, rows=(r0, r1), col_rot=(r0, world * block))
        assert torch.equal(rpr, rp) and torch.equal(dvr, dv)
        for i in range(r1 - r0):
            a, b = cl[rp[i]:rp[i + 1]].long(), clr[rp[i]:rp[i + 1]].long()
            assert torch.equal(torch.sort((a - r0) % (world * block))[0], b)
            assert bool((b[1:] >= b[:-1]).all())
