"""Import-only stub (test infra)."""


class NodePropPredDataset:
    def __init__(self, *a, **k):
        raise RuntimeError("ogb stub: no datasets on this box")


class PygNodePropPredDataset(NodePropPredDataset):
    pass
