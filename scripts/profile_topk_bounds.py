"""ncu workload: LJ-shape top-k (k=500) with the non-opt driver (bounds), one wave of 16 queries (development script)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fora_b200 as fb
n, m = 4847571, 68993773
src, dst = fb.synth_edges(n, m, 42)
E = fb.Engine(0, seed=2026, slots=16)
E.build_graph_from_edges(n, m, src, dst, with_in=False)
q = np.random.default_rng(43).integers(0, n, 1000).astype(np.int32)
E.configure("fora", 0.5, k=500, opt=0)
*_, st, tm = E.topk_batch("fora", q[:16], 500)
print(tm)
