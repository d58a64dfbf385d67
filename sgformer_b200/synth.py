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


def make_rmat_graph(n: int, e_stored: int, seed: int = 0, device="cpu", a: float = 0.57, b: float = 0.19, c: float = 0.19) -> torch.Tensor:
    """Power-law (R-MAT, a=.57 b=c=.19) variant of `make_graph` (SURVEY.md §8d secondary workload): skewed degrees with hub
    rows, same symmetrise / self-loop post-processing.  Node ids are drawn in [0, 2^ceil(log2 n)) and folded into [0, n)."""
    g = torch.Generator(device=device).manual_seed(seed)
    bits = max(1, (n - 1).bit_length())
    src = torch.zeros(e_stored, dtype=torch.int64, device=device)
    dst = torch.zeros(e_stored, dtype=torch.int64, device=device)
    for _ in range(bits):
        r = torch.rand(e_stored, generator=g, device=device)
        src = src * 2 + (r >= a + b).long()                       # quadrants c, d set the source bit
        dst = dst * 2 + (((r >= a) & (r < a + b)) | (r >= a + b + c)).long()   # quadrants b, d set the target bit
    src, dst = src % n, dst % n
    key = torch.unique(torch.cat([src * n + dst, dst * n + src]))
    row, col = key // n, key % n
    keep = row != col
    row, col = row[keep], col[keep]
    loops = torch.arange(n, device=device)
    return torch.stack([torch.cat([row, loops]), torch.cat([col, loops])]).contiguous()
