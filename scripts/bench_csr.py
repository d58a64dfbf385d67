"""Micro-benchmark of the graph build at the products shape: sgf_csr_build (count / scan / fill / sort / dinv) + sgf_edge_symmetry.
    SGF_CSR_FILL_WINDOW_MB=<mb> python scripts/bench_csr.py
CUDA events, 3 warm-ups, 5 timed builds; prints one line."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgformer_b200 import kernels as K  # noqa: E402
from sgformer_b200.synth import make_graph  # noqa: E402

n, e = 2_449_029, 61_859_140
ei = make_graph(n, e, seed=0, device="cuda")
torch.cuda.synchronize()


def timed(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


t_build = timed(lambda: K.csr_build(ei, n, False, 0, True))
t_sym = timed(lambda: K.edge_symmetry(ei, n))
print(f"buckets={os.environ.get('SGF_CSR_BUCKETS', 'default')} window_mb={os.environ.get('SGF_CSR_FILL_WINDOW_MB', 'default')} nnz={ei.shape[1]} csr_build {t_build:.3f} ms  edge_symmetry {t_sym:.3f} ms")
