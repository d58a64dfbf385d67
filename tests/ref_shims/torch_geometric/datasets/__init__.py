"""Dataset shims (test infra).  `Planetoid` returns a SEEDED SYNTHETIC graph with the named dataset's shape (no network,
no dataset files on this box) so that the reference's medium/main.py runs end to end as plumbing (BASELINE config 1:
Cora-shaped, 2708 nodes / 1433 features / 5278 stored edges / 7 classes)."""
import torch

_SHAPES = {"cora": (2708, 1433, 5278, 7), "citeseer": (3327, 3703, 4552, 6), "pubmed": (19717, 500, 44324, 3)}


class _Data:
    pass


class Planetoid:
    def __init__(self, root=None, name="cora", transform=None, **kw):
        n, d, e, c = _SHAPES[name.lower()]
        g = torch.Generator().manual_seed(0)
        src = torch.randint(0, n, (e,), generator=g)
        dst = torch.randint(0, n, (e,), generator=g)
        data = _Data()
        data.x = (torch.rand(n, d, generator=g) < 0.01).float()
        data.edge_index = torch.cat([torch.stack([src, dst]), torch.stack([dst, src])], 1)
        data.y = torch.randint(0, c, (n,), generator=g)
        data.num_nodes = n
        idx = torch.randperm(n, generator=g)
        for nm, sl in (("train_mask", idx[:140]), ("val_mask", idx[140:640]), ("test_mask", idx[640:1640])):
            m = torch.zeros(n, dtype=torch.bool)
            m[sl] = True
            setattr(data, nm, m)
        self._data = data

    def __getitem__(self, i):
        assert i == 0
        return self._data

    def __len__(self):
        return 1


class _NoData:
    def __init__(self, *a, **k):
        raise RuntimeError("dataset stub: no datasets on this box (synthetic graphs are used)")


class Amazon(_NoData):
    pass


class Coauthor(_NoData):
    pass
