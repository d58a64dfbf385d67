"""Device versions of the torch_geometric.utils functions the reference drivers call around the model:

    to_undirected      large/main.py:76, medium/main.py:94
    remove_self_loops  large/main.py:78, large/main-batch.py:97
    add_self_loops     large/main.py:79, large/main-batch.py:98
    subgraph           large/main-batch.py:139, large/eval.py:89   (relabel_nodes=True)

Same call signatures and return conventions as torch_geometric 1.7.2 for the argument combinations the reference uses (no edge
attributes); results are bit-identical (tests/test_gpu_kernels.py against oracle/np_ref.py).  A maintainer swaps the import:

    from sgformer_b200.pyg_utils import to_undirected, remove_self_loops, add_self_loops, subgraph

Inputs must live on a CUDA device: there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Tuple

from torch import Tensor

from . import kernels as K


def _num_nodes(edge_index: Tensor, num_nodes: Optional[int]) -> int:
    if num_nodes is not None:
        return int(num_nodes)
    return int(edge_index.max().item()) + 1 if edge_index.numel() else 0


def _no_attr(edge_attr):
    if edge_attr is not None:
        raise NotImplementedError("sgformer_b200.pyg_utils: edge attributes are not used by the SGFormer drivers")


def to_undirected(edge_index: Tensor, edge_attr=None, num_nodes: Optional[int] = None) -> Tensor:
    if isinstance(edge_attr, int):          # torch_geometric 1.7.2 accepts to_undirected(edge_index, num_nodes)
        num_nodes, edge_attr = edge_attr, None
    _no_attr(edge_attr)
    return K.to_undirected(edge_index, _num_nodes(edge_index, num_nodes))


def remove_self_loops(edge_index: Tensor, edge_attr=None) -> Tuple[Tensor, None]:
    _no_attr(edge_attr)
    return K.remove_self_loops(edge_index), None


def add_self_loops(edge_index: Tensor, edge_attr=None, fill_value=None, num_nodes: Optional[int] = None) -> Tuple[Tensor, None]:
    _no_attr(edge_attr)
    return K.add_self_loops(edge_index, _num_nodes(edge_index, num_nodes)), None


def subgraph(subset: Tensor, edge_index: Tensor, edge_attr=None, relabel_nodes: bool = False, num_nodes: Optional[int] = None):
    _no_attr(edge_attr)
    if not relabel_nodes:
        raise NotImplementedError("sgformer_b200.pyg_utils.subgraph: the reference only calls it with relabel_nodes=True")
    return K.subgraph(edge_index, _num_nodes(edge_index, num_nodes), subset), None
