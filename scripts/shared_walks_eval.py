"""Shared walks (opt-in per-wave walk pool) against private walks on the bench workload (development script).
usage: python scripts/shared_walks_eval.py [slots] [nq]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fora_b200 as fb
S = int(sys.argv[1]) if len(sys.argv) > 1 else 48
NQ = int(sys.argv[2]) if len(sys.argv) > 2 else 96
n, m = 4847571, 68993773
src, dst = fb.synth_edges(n, m, 42)
op, oc, _, _ = fb.csr_from_edges(n, src, dst, with_in=False)
q = np.random.default_rng(43).integers(0, n, 1000).astype(np.int32)
E = fb.Engine(0, seed=2026, slots=S)
E.upload_graph(n, m, op, oc)
E.configure("fora", 0.5, opt=1, balanced=1)
for shared in (0, 1, 0, 1):
    E.set_shared_walks(bool(shared))
    E.query_batch("fora", q[:S], want_ppr=False)
    _, st, tm = E.query_batch("fora", q[S:S + NQ], want_ppr=False)
    walks = sum(s["n_walks"] for s in st) / NQ
    print("shared %d slots %d: %7.1f q/s | push %.3f ms/q | walk phase %.3f ms/q (walk kernels %.3f) | walks per query %.1f M, launches %d" % (
        shared, S, NQ / (tm["total_ms"] * 1e-3), tm["push_ms"] / NQ, tm["walk_ms"] / NQ, tm["walk_kernel_ms"] / NQ, walks / 1e6, tm["kernel_launches"]), flush=True)
