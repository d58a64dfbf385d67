"""Drop-in for the reference's 100M/ours.py (`--method ours`, 100M/parse.py:4-8): the large structure plus the
alpha-residual of the attention branch and the free function `full_attention_conv`."""
import torch.nn.functional as F

from .large import SGFormer as _LargeSGFormer
from .modules import (GraphConvBase, GraphConvLayerBase, TransConvBase, TransConvLayerBase, full_attention_conv)

__all__ = ["full_attention_conv", "GraphConvLayer", "GraphConv", "TransConvLayer", "TransConv", "SGFormer"]


class GraphConvLayer(GraphConvLayerBase):
    """100M/ours.py:56-88"""


class GraphConv(GraphConvBase):
    """100M/ours.py:91-152"""
    variant = "100M"

    def _layer_cls(self):
        return GraphConvLayer


class TransConvLayer(TransConvLayerBase):
    """100M/ours.py:155-195"""

    def forward(self, query_input, source_input, edge_index=None, output_attn=False):
        return self._attend(query_input, source_input, output_attn)


class TransConv(TransConvBase):
    """100M/ours.py:198-289 (residual = alpha*x + (1-alpha)*prev, :264)"""
    variant = "100M"

    def __init__(self, in_channels, hidden_channels, num_layers=2, num_heads=1, alpha=0.5, dropout=0.5, use_bn=True,
                 use_residual=True, use_weight=True, use_act=True):
        super().__init__()
        self._build(in_channels, hidden_channels, num_layers, num_heads, use_weight, TransConvLayer)
        self.dropout = dropout
        self.activation = F.relu
        self.use_bn = use_bn
        self.use_residual = use_residual
        self.alpha = alpha
        self.use_act = use_act

    def forward(self, x, edge_index=None):
        return self._run(x)

    def get_attentions(self, x):
        return self._attentions(x, with_act=False)


class SGFormer(_LargeSGFormer):
    """100M/ours.py:292-380"""
    variant = "100M"

    def __init__(self, in_channels, hidden_channels, out_channels,
                 trans_num_layers=1, trans_num_heads=1, trans_dropout=0.5,
                 gnn_num_layers=1, gnn_dropout=0.5, gnn_use_weight=True, gnn_use_init=False, gnn_use_bn=True,
                 gnn_use_residual=True, gnn_use_act=True, alpha=0.5,
                 trans_use_bn=True, trans_use_residual=True, trans_use_weight=True, trans_use_act=True,
                 use_graph=True, graph_weight=0.8, aggregate="add"):
        # build through nn.Module.__init__ to keep the reference's submodule registration order
        super(_LargeSGFormer, self).__init__()
        self.trans_conv = TransConv(in_channels, hidden_channels, trans_num_layers, trans_num_heads, alpha, trans_dropout,
                                    trans_use_bn, trans_use_residual, trans_use_weight, trans_use_act)
        self.graph_conv = GraphConv(in_channels, hidden_channels, gnn_num_layers, gnn_dropout, gnn_use_bn,
                                    gnn_use_residual, gnn_use_weight, gnn_use_init, gnn_use_act)
        self.use_graph = use_graph
        self.graph_weight = graph_weight
        self.alpha = alpha
        self.aggregate = aggregate
        self._finish_init(hidden_channels, out_channels, aggregate)
        self.params1 = list(self.trans_conv.parameters())
        self.params2 = list(self.graph_conv.parameters()) if self.graph_conv is not None else []
        self.params2.extend(list(self.fc.parameters()))
        self._io = (in_channels, hidden_channels, out_channels)
