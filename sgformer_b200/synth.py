"""Seeded synthetic graphs shaped like the reference's datasets: src,dst ~ U{0..N-1} i.i.d. ->
to_undirected (symmetrise + coalesce) -> remove_self_loops -> add_self_loops, mirroring large/main.py:75-79.
Runs on the device of `device` (default CPU so the CPU oracle and the GPU see identical data)."""
import torch

# name -> (N, d_in, E_stored, classes, hidden, gnn layers, use_init)   [table at the top of SURVEY.md §8]
SHAPES = {
    "cora": (2708, 1433, 5278, 7, 64, 4, False),
    "arxiv": (169343, 128, 1166243, 40, 256, 3, False),
    "products": (2449029, 100, 61859140, 47, 256, 3, True),
    "pokec": (1632803, 65, 30622564, 2, 64, 2, True),
    "papers100M": (111059956, 128, 1615685872, 172, 256, 3, True),
}


def make_graph(n: int, e_stored: int, seed: int = 0, device="cpu") -> torch.Tensor:
    g = torch.Generator(device=device).manual_seed(seed)
    src = torch.randint(0, n, (e_stored,), generator=g, device=device)
    dst = torch.randint(0, n, (e_stored,), generator=g, device=device)
    key = torch.cat([src * n + dst, dst * n + src])
    del src, dst
    key = torch.unique(key)                       # coalesced, sorted by (row, col)
    row, col = key // n, key % n
    del key
    keep = row != col
    row, col = row[keep], col[keep]
    loops = torch.arange(n, device=device)
    return torch.stack([torch.cat([row, loops]), torch.cat([col, loops])]).contiguous()
