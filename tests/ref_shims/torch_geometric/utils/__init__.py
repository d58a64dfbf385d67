"""PyG utility semantics restated (test infra). Call sites: large/ours.py:28,
large/main.py:76-79, large/main-batch.py:139, medium/main.py:94."""
import torch


def _num_nodes(edge_index, num_nodes):
    if num_nodes is not None:
        return int(num_nodes)
    return int(edge_index.max()) + 1 if edge_index.numel() else 0


def degree(index, num_nodes=None, dtype=None):
    n = _num_nodes(index, num_nodes)
    out = torch.zeros(n, dtype=dtype or torch.get_default_dtype(), device=index.device)
    return out.scatter_add_(0, index, torch.ones(index.numel(), dtype=out.dtype, device=index.device))


def remove_self_loops(edge_index, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], (None if edge_attr is None else edge_attr[keep])


def add_self_loops(edge_index, edge_attr=None, fill_value=1.0, num_nodes=None):
    n = _num_nodes(edge_index, num_nodes)
    loops = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    ei = torch.cat([edge_index, torch.stack([loops, loops])], dim=1)
    if edge_attr is not None:
        edge_attr = torch.cat([edge_attr, edge_attr.new_full((n,) + tuple(edge_attr.shape[1:]), fill_value)])
    return ei, edge_attr


def add_remaining_self_loops(edge_index, edge_attr=None, fill_value=1.0, num_nodes=None):
    n = _num_nodes(edge_index, num_nodes)
    keep = edge_index[0] != edge_index[1]
    loops = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    ei = torch.cat([edge_index[:, keep], torch.stack([loops, loops])], dim=1)
    if edge_attr is not None:
        loop_attr = edge_attr.new_full((n,) + tuple(edge_attr.shape[1:]), fill_value)
        inv = ~keep
        loop_attr[edge_index[0][inv]] = edge_attr[inv]
        edge_attr = torch.cat([edge_attr[keep], loop_attr])
    return ei, edge_attr


def coalesce(edge_index, num_nodes=None):
    n = _num_nodes(edge_index, num_nodes)
    key = torch.unique(edge_index[0] * n + edge_index[1])  # sorted, deduplicated
    return torch.stack([key // n, key % n])


def to_undirected(edge_index, edge_attr=None, num_nodes=None):
    assert edge_attr is None or isinstance(edge_attr, int)
    if isinstance(edge_attr, int):
        num_nodes = edge_attr
    both = torch.cat([edge_index, edge_index.flip(0)], dim=1)
    return coalesce(both, num_nodes)


def subgraph(subset, edge_index, edge_attr=None, relabel_nodes=False, num_nodes=None,
             return_edge_mask=False):
    n = _num_nodes(edge_index, num_nodes)
    if subset.dtype == torch.bool:
        node_mask = subset
        subset = node_mask.nonzero().view(-1)
    else:
        node_mask = torch.zeros(n, dtype=torch.bool, device=edge_index.device)
        node_mask[subset] = True
    edge_mask = node_mask[edge_index[0]] & node_mask[edge_index[1]]
    ei = edge_index[:, edge_mask]
    if relabel_nodes:
        relabel = torch.zeros(n, dtype=torch.long, device=edge_index.device)
        relabel[subset] = torch.arange(subset.numel(), device=edge_index.device)
        ei = relabel[ei]
    ea = None if edge_attr is None else edge_attr[edge_mask]
    if return_edge_mask:
        return ei, ea, edge_mask
    return ei, ea


def k_hop_subgraph(*a, **k):
    raise NotImplementedError("stub")


def to_dense_adj(edge_index, batch=None, edge_attr=None, max_num_nodes=None):
    n = _num_nodes(edge_index, max_num_nodes)
    adj = torch.zeros(1, n, n, device=edge_index.device)
    adj[0].index_put_((edge_index[0], edge_index[1]),
                      torch.ones(edge_index.shape[1], device=edge_index.device), accumulate=True)
    return adj
