"""Graph preprocessing around the model: the numpy oracle (oracle/np_ref.py) against the torch restatement of
torch_geometric 1.7.2 that the unmodified reference drivers run through in this container (tests/ref_shims).  torch_geometric itself
is not installable here: both sides restate its documented semantics independently — parity with the real package
is UNPINNED, the call sites (large/main.py:75-79, large/main-batch.py:97-98, medium/main.py:94) fix what must be computed."""
import os
import sys

import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "ref_shims"))

from torch_geometric.utils import add_self_loops, remove_self_loops, to_undirected  # noqa: E402  (the shim)

from oracle import np_ref  # noqa: E402


@settings(max_examples=60, deadline=None)
@given(n=st.integers(1, 40), e=st.integers(0, 200), seed=st.integers(0, 10 ** 6), loops=st.integers(0, 10))
def test_preprocessing_oracle_matches_shim(n, e, seed, loops):
    g = torch.Generator().manual_seed(seed)
    ei = torch.stack([torch.randint(0, n, (e,), generator=g), torch.randint(0, n, (e,), generator=g)])
    if e:
        k = min(loops, e)
        ei[1, :k] = ei[0, :k]
    a = ei.numpy()
    if e:
        assert np.array_equal(np_ref.to_undirected(a, n), to_undirected(ei, num_nodes=n).numpy())
    assert np.array_equal(np_ref.remove_self_loops(a), remove_self_loops(ei)[0].numpy())
    assert np.array_equal(np_ref.add_self_loops(a, n), add_self_loops(ei, num_nodes=n)[0].numpy())


def test_preprocessing_known_answer():
    ei = np.array([[0, 2, 2, 1, 3, 0], [1, 2, 0, 0, 3, 1]])
    assert np_ref.to_undirected(ei, 4).tolist() == [[0, 0, 1, 2, 2, 3], [1, 2, 0, 0, 2, 3]]
    assert np_ref.remove_self_loops(ei).tolist() == [[0, 2, 1, 0], [1, 0, 0, 1]]
    assert np_ref.add_self_loops(ei[:, :2], 3).tolist() == [[0, 2, 0, 1, 2], [1, 2, 0, 1, 2]]
