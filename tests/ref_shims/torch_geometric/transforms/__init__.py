"""Import-only stubs (test infra)."""


class NormalizeFeatures:
    def __call__(self, data):
        return data


class ToUndirected:
    def __call__(self, data):
        return data
