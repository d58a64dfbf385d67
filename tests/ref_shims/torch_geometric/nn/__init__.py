"""GCNConv restated from PyG's published semantics; the rest are import-only stubs (test infra).
Call site: medium/models.py:14-63."""
import torch.nn as nn
from .conv.gcn_conv import GCNConv, gcn_norm  # noqa: F401


class _Stub(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("torch_geometric.nn stub (baseline model, out of scope)")


class MessagePassing(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()


class SGConv(_Stub):
    pass


class GATConv(_Stub):
    pass


class JumpingKnowledge(_Stub):
    pass


class APPNP(_Stub):
    pass
