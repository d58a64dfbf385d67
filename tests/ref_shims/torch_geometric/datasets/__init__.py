"""Import-only stubs (test infra)."""


class _NoData:
    def __init__(self, *a, **k):
        raise RuntimeError("dataset stub: no datasets on this box (synthetic graphs are used)")


class Planetoid(_NoData):
    pass


class Amazon(_NoData):
    pass


class Coauthor(_NoData):
    pass
