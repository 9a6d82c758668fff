"""Workload for the ncu captures under profiles/: LJ-shape, FORA eps=0.5 --balanced --opt, one wave of queries."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fora_b200 as fb
slots = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nq = int(sys.argv[2]) if len(sys.argv) > 2 else slots
n, m = 4847571, 68993773
src, dst = fb.synth_edges(n, m, 42)
op, oc, _, _ = fb.csr_from_edges(n, src, dst, with_in=False)
E = fb.Engine(0, seed=2026, slots=slots)
E.upload_graph(n, m, op, oc)
E.configure("fora", 0.5, opt=1, balanced=1)
q = np.random.default_rng(43).integers(0, n, 1000).astype(np.int32)
_, stats, tm = E.query_batch("fora", q[:nq], want_ppr=False)
print(tm)
print("edges", sum(s["edges_pushed"] for s in stats), "verts", sum(s["vertices_pushed"] for s in stats), "walks", sum(s["n_walks"] for s in stats), "hops", sum(s["walk_hops"] for s in stats), "rounds", [s["push_rounds"] for s in stats])
