"""PyG GCNConv / gcn_norm semantics restated (test infra).

gcn_norm (improved=False, add_self_loops=True): add_remaining_self_loops with fill 1; deg = scatter_add of
edge weights at `col`; w_e = deg^-1/2[row] * w_e * deg^-1/2[col] (inf -> 0).
GCNConv: x' = x W (weight [in,out], glorot), propagate: out[col] += w_e * x'[row]; + bias (zeros init)."""
import math
import torch
import torch.nn as nn
from ...utils import add_remaining_self_loops


def gcn_norm(edge_index, edge_weight=None, num_nodes=None, improved=False, add_self_loops=True,
             dtype=None):
    fill = 2.0 if improved else 1.0
    n = int(num_nodes) if num_nodes is not None else int(edge_index.max()) + 1
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.shape[1], dtype=dtype or torch.float32,
                                 device=edge_index.device)
    if add_self_loops:
        edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, fill, n)
    row, col = edge_index[0], edge_index[1]
    deg = torch.zeros(n, dtype=edge_weight.dtype, device=edge_index.device).scatter_add_(0, col, edge_weight)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    return edge_index, dis[row] * edge_weight * dis[col]


class _Lin(nn.Module):
    """PyG>=2 stores GCNConv's weight as `lin.weight` of shape [out, in] (bias-free Linear)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin))

    def forward(self, x):
        return x @ self.weight.t()


class GCNConv(nn.Module):
    def __init__(self, in_channels, out_channels, improved=False, cached=False,
                 add_self_loops=True, normalize=True, bias=True, **kw):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached = improved, cached
        self.add_self_loops, self.normalize = add_self_loops, normalize
        self.lin = _Lin(in_channels, out_channels)
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        a = math.sqrt(6.0 / (self.in_channels + self.out_channels))
        nn.init.uniform_(self.lin.weight, -a, a)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x, edge_index, edge_weight=None):
        n = x.shape[0]
        if self.normalize:
            edge_index, edge_weight = gcn_norm(edge_index, edge_weight, n, self.improved,
                                               self.add_self_loops, x.dtype)
        elif edge_weight is None:
            edge_weight = torch.ones(edge_index.shape[1], dtype=x.dtype, device=x.device)
        x = self.lin(x)
        out = torch.zeros(n, x.shape[1], dtype=x.dtype, device=x.device)
        out = out.index_add(0, edge_index[1], x[edge_index[0]] * edge_weight.unsqueeze(-1))
        if self.bias is not None:
            out = out + self.bias
        return out
