"""Shim of the torch_geometric import surface the reference needs (test infra)."""
from . import utils, nn, datasets, transforms, data, loader  # noqa: F401


def seed_everything(seed):
    import random
    import numpy as np
    import torch
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
