"""ncu workload: one balanced FORA wave with shared walks on the LJ-shape graph (development script)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fora_b200 as fb
slots = int(sys.argv[1]) if len(sys.argv) > 1 else 48
n, m = 4847571, 68993773
src, dst = fb.synth_edges(n, m, 42)
op, oc, _, _ = fb.csr_from_edges(n, src, dst, with_in=False)
E = fb.Engine(0, seed=2026, slots=slots)
E.upload_graph(n, m, op, oc)
E.configure("fora", 0.5, opt=1, balanced=1)
E.set_shared_walks(True)
q = np.random.default_rng(43).integers(0, n, 1000).astype(np.int32)
_, stats, tm = E.query_batch("fora", q[:slots], want_ppr=False)
print(tm)
