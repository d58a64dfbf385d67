"""Host logic of the halo-buffer reuse fence (sgformer_b200.dist.Comm._fence_reuse): a 4-byte all-reduce is issued before a
symmetric buffer is reused iff no collective went through the Comm since its last use.  No GPU, no process group: the collective
is replaced by a counter."""
import torch

from sgformer_b200 import dist as D


class _State:
    fence = torch.zeros(1)


def _comm(monkeypatch):
    calls = []
    monkeypatch.setattr(D.dist, "all_reduce", lambda t, group=None: calls.append(t))
    c = D.Comm(None)
    c.active, c.world, c.rank, c.group = True, 2, 0, None      # pretend to be one rank of two
    return c, calls


def test_fence_only_without_an_intervening_collective(monkeypatch):
    c, calls = _comm(monkeypatch)
    st = _State()
    c._fence_reuse(st, 0)                 # first use of buffer 0: nothing to order
    c._last_use[0] = c._coll_seq
    assert calls == []
    c._fence_reuse(st, 0)                 # reused with no collective in between: fence
    assert len(calls) == 1 and calls[0] is st.fence
    c._last_use[0] = c._coll_seq
    c.allreduce_(torch.zeros(3))          # a step collective (C1 / C3 / C5) ...
    n = len(calls)
    c._fence_reuse(st, 0)                 # ... orders the reuse by itself
    assert len(calls) == n
    c._last_use[0] = c._coll_seq
    c._fence_reuse(st, 1)                 # another buffer, never used
    assert len(calls) == n
    c._fence_reuse(st, 0)                 # and buffer 0 again without a collective
    assert len(calls) == n + 1


def test_every_rank_takes_the_same_branch(monkeypatch):
    """The decision depends only on the schedule (counters), never on data: two Comms driven by the same call sequence agree."""
    seqs = []
    for _ in range(2):
        c, calls = _comm(monkeypatch)
        st = _State()
        trace = []
        for step in range(3):
            for k in (0, 1):
                before = len(calls)
                c._fence_reuse(st, k)
                trace.append(len(calls) - before)
                c._last_use[k] = c._coll_seq
            if step == 1:
                c.allreduce_(torch.zeros(2))
        seqs.append(trace)
    assert seqs[0] == seqs[1]
    assert seqs[0][:2] == [0, 0] and seqs[0][2:4] == [1, 0]      # second step: buffer 0 fenced; the fence itself orders buffer 1
