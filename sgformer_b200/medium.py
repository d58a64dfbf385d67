"""Drop-in for the reference's medium/ours.py (`--method ours`, medium/parse.py:97-104) plus a native `GCN` backbone
with the semantics of medium/models.py:14-63 (PyG GCNConv stack) for users without torch_geometric."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import engine as E
from . import functional as Fn
from .config import make_config
from .dist import SINGLE
from .graph import get_graph
from .modules import SGFormerBase, TransConvBase, TransConvLayerBase, _Base, full_attention_conv

__all__ = ["full_attention_conv", "TransConvLayer", "TransConv", "SGFormer", "GCN", "GCNConv"]


class TransConvLayer(TransConvLayerBase):
    """medium/ours.py:49-100"""

    def forward(self, query_input, source_input, edge_index=None, edge_weight=None, output_attn=False):
        return self._attend(query_input, source_input, output_attn)


class TransConv(TransConvBase):
    """medium/ours.py:103-177 (takes the dataset object; residual = alpha*x + (1-alpha)*prev, :152)"""
    variant = "medium"

    def __init__(self, in_channels, hidden_channels, num_layers=2, num_heads=1, alpha=0.5, dropout=0.5, use_bn=True,
                 use_residual=True, use_weight=True, use_act=False):
        super().__init__()
        self._build(in_channels, hidden_channels, num_layers, num_heads, use_weight, TransConvLayer)
        self.dropout = dropout
        self.activation = F.relu
        self.use_bn = use_bn
        self.residual = use_residual
        self.alpha = alpha
        self.use_act = use_act

    def forward(self, data):
        return self._run(data.graph['node_feat'])

    def get_attentions(self, x):
        return self._attentions(x, with_act=False)


class _Lin(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin))


class GCNConv(nn.Module):
    """Parameter container with PyG>=2 GCNConv's names (`lin.weight` [out,in], `bias`) and inits (glorot / zeros)."""

    def __init__(self, in_channels, out_channels, cached=False, **kw):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _Lin(in_channels, out_channels)
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        a = math.sqrt(6.0 / (self.in_channels + self.out_channels))
        nn.init.uniform_(self.lin.weight, -a, a)
        nn.init.zeros_(self.bias)


class GCN(_Base):
    """models.GCN (medium/models.py:14-63): `num_layers` GCNConv layers, BN/ReLU/dropout between them."""

    def __init__(self, in_channels, hidden_channels, out_channels, num_layers=2, dropout=0.5, save_mem=True, use_bn=True):
        super().__init__()
        self.convs = nn.ModuleList()
        self.convs.append(GCNConv(in_channels, hidden_channels, cached=not save_mem))
        self.bns = nn.ModuleList()
        self.bns.append(nn.BatchNorm1d(hidden_channels))
        for _ in range(num_layers - 2):
            self.convs.append(GCNConv(hidden_channels, hidden_channels, cached=not save_mem))
            self.bns.append(nn.BatchNorm1d(hidden_channels))
        self.convs.append(GCNConv(hidden_channels, out_channels, cached=not save_mem))
        self.dropout = dropout
        self.activation = F.relu
        self.use_bn = use_bn

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()
        for bn in self.bns:
            bn.reset_parameters()

    def forward(self, data):
        x, edge_index = data.graph['node_feat'], data.graph['edge_index']
        if 'edge_weight' in data.graph and data.graph['edge_weight'] is not None:
            raise NotImplementedError("sgformer_b200.GCN: edge_weight is not supported")
        if not x.is_cuda:
            raise RuntimeError("sgformer_b200.GCN needs CUDA tensors (no CPU fallback)")
        names, tensors = _gcn_flat(self, "gnn.")
        cfg = make_config("medium", x.shape[1], self.convs[0].out_channels, self.convs[-1].out_channels,
                          gcn_num_layers=len(self.convs), gcn_dropout=self.dropout, gcn_use_bn=self.use_bn)
        graph = get_graph(edge_index, x.shape[0], 1)
        return Fn.GraphBranchFn.apply(x, graph, cfg, E.precision(self.precision), self.training, "gcn", "gnn.", names,
                                      *tensors)


def _is_gcn_like(gnn) -> bool:
    """A models.GCN-shaped module (ours, or the reference's over PyG GCNConv) whose layers we can run natively."""
    try:
        convs, bns = gnn.convs, gnn.bns
        if len(convs) < 1 or not hasattr(gnn, "dropout") or not hasattr(gnn, "use_bn"):
            return False
        for c in convs:
            w = c.lin.weight if hasattr(c, "lin") else c.weight
            if w.dim() != 2 or getattr(c, "bias", None) is None:
                return False
            if getattr(c, "improved", False) or not getattr(c, "normalize", True) or not getattr(c, "add_self_loops", True):
                return False
        return len(bns) >= len(convs) - 1
    except AttributeError:
        return False


def _gcn_out_dim(gnn) -> int:
    c = gnn.convs[-1]
    return c.lin.weight.shape[0] if hasattr(c, "lin") else c.weight.shape[1]


def _gcn_flat(gnn, prefix):
    """(names, tensors) with PyG>=2 naming; PyG 1.x stores GCNConv.weight as [in,out] -> transposed view (autograd
    carries the gradient back through the transpose)."""
    names, tensors = [], []
    for i, c in enumerate(gnn.convs):
        if hasattr(c, "lin"):
            w = c.lin.weight
        else:
            w = c.weight.t().contiguous()
        names.append(f"{prefix}convs.{i}.lin.weight"); tensors.append(w)
        names.append(f"{prefix}convs.{i}.bias"); tensors.append(c.bias)
    for i, bn in enumerate(gnn.bns):
        for nm in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
            names.append(f"{prefix}bns.{i}.{nm}"); tensors.append(getattr(bn, nm))
    return tuple(names), tensors


class SGFormer(SGFormerBase):
    """medium/ours.py:179-223: attention branch + an injected GNN (`gnn=`)."""
    variant = "medium"
    _self_loop_mode = 1

    def __init__(self, in_channels, hidden_channels, out_channels, num_layers=2, num_heads=1, alpha=0.5, dropout=0.5,
                 use_bn=True, use_residual=True, use_weight=True, use_graph=True, use_act=False, graph_weight=0.8,
                 gnn=None, aggregate='add'):
        super().__init__()
        # medium/ours.py:183 does not forward use_act to TransConv
        self.trans_conv = TransConv(in_channels, hidden_channels, num_layers, num_heads, alpha, dropout, use_bn,
                                    use_residual, use_weight)
        self.gnn = gnn
        self.use_graph = use_graph
        self.graph_weight = graph_weight
        self.use_act = use_act
        self.aggregate = aggregate
        self._finish_init(hidden_channels, out_channels, aggregate)
        self.params1 = list(self.trans_conv.parameters())
        self.params2 = list(self.gnn.parameters()) if self.gnn is not None else []
        self.params2.extend(list(self.fc.parameters()))
        self._io = (in_channels, hidden_channels, out_channels)

    def _cfg(self, gcn_layers=2, gcn_dropout=0.5, gcn_use_bn=True) -> dict:
        d, h, c = self._io
        t = self.trans_conv
        _, _, tnl, tnh = t._dims
        return make_config("medium", d, h, c, trans_num_layers=tnl, num_heads=tnh, trans_dropout=t.dropout,
                           trans_use_bn=t.use_bn, trans_use_residual=t.residual,
                           trans_use_weight=t.convs[0].use_weight if tnl else True, trans_use_act=False, alpha=t.alpha,
                           use_graph=bool(self.use_graph), graph_weight=float(self.graph_weight),
                           aggregate=self.aggregate, gcn_num_layers=gcn_layers, gcn_dropout=gcn_dropout,
                           gcn_use_bn=gcn_use_bn)

    def forward(self, data):
        x, edge_index = data.graph['node_feat'], data.graph['edge_index']
        if not x.is_cuda:
            raise RuntimeError("sgformer_b200.SGFormer (medium) needs the dataset on a CUDA device (no CPU fallback)")
        prec = E.precision(self.precision)
        fused = bool(self.use_graph) and self.gnn is not None and _is_gcn_like(self.gnn) and \
            data.graph.get('edge_weight', None) is None and _gcn_out_dim(self.gnn) == self._io[1]
        tn, tt = self.trans_conv._flat("trans_conv.")
        fn, ft = self.fc._parameters.keys(), list(self.fc._parameters.values())
        names = list(tn) + ["fc." + k for k in fn]
        tensors = list(tt) + ft
        if not self.use_graph:
            return Fn.SGFormerFn.apply(x, None, self._cfg(), prec, self.training, self._comm, tuple(names), *tensors)
        if fused:
            gn, gt = _gcn_flat(self.gnn, "gnn.")
            cfg = self._cfg(len(self.gnn.convs), float(self.gnn.dropout), bool(self.gnn.use_bn))
            comm = self._comm
            graph = get_graph(edge_index, comm.n_global, 1, rows=comm.rows, col_rot=comm.col_rot) if comm.active else get_graph(edge_index, x.shape[0], 1)
            return Fn.SGFormerFn.apply(x, graph, cfg, prec, self.training, comm, tuple(names) + tuple(gn), *tensors, *gt)
        # foreign GNN module: run it as given, mix + fc on the GPU kernels
        x1 = Fn.TransConvFn.apply(x, self.trans_conv._cfg(), prec, self.training, tn, *tt)
        x2 = self.gnn(data)
        return Fn.HeadFn.apply(x1, x2, self._cfg(), prec, tuple("fc." + k for k in fn), *ft)

    def reset_parameters(self):
        self.trans_conv.reset_parameters()
        if self.use_graph:
            self.gnn.reset_parameters()
