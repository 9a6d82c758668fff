"""Round-2 experiment (a): the shipped push kernel at small slot counts (sub-waves whose residue vectors are L2-resident).

usage: python scripts/r2_subwave.py [slot counts ...]   -> one line per slot count (push ms/query, G edges/s, walk)
Env FORA_* knobs pass through to the library.  With R2_TRACE=1 the per-level trace of one NON-balanced launch is printed too.
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRACE = os.environ.get("R2_TRACE") == "1"
if TRACE:
    os.environ["FORA_PUSH_TRACE"] = "1"
import fora_b200 as fb  # noqa: E402

n, m = 4847571, 68993773
t0 = time.time()
src, dst = fb.synth_edges(n, m, 42)
op, oc, _, _ = fb.csr_from_edges(n, src, dst, with_in=False)
del src, dst
queries = np.random.default_rng(43).integers(0, n, 1000).astype(np.int32)
print("graph ready in %.1f s" % (time.time() - t0), flush=True)
slot_list = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 6, 8, 12, 16, 24, 48]
for slots in slot_list:
    nq = max(2 * slots, 24)
    nq -= nq % slots
    E = fb.Engine(0, seed=2026, slots=slots)
    try:
        E.upload_graph(n, m, op, oc)
        for bal in (1, 0):
            E.configure("fora", 0.5, opt=1, balanced=bal)
            E.query_batch("fora", queries[:slots], want_ppr=False)
            _, stats, tm = E.query_batch("fora", queries[slots:slots + nq], want_ppr=False)
            ed = sum(s["edges_pushed"] for s in stats) / nq
            lv = sum(s["push_levels"] for s in stats) / nq
            hops = sum(s["walk_hops"] for s in stats) / nq
            print("slots %2d balanced %d: %6.1f q/s | push kernel %.3f ms/q (%.1f G edges/s, %.1fM edges, %.0f levels) push phase %.3f | walk kernel %.3f ms/q (%.1f G hops/s)" % (
                slots, bal, nq / (tm["total_ms"] * 1e-3), tm["push_kernel_ms"] / nq, ed / (tm["push_kernel_ms"] / nq) / 1e6, ed / 1e6, lv,
                tm["push_ms"] / nq, tm["walk_kernel_ms"] / nq, hops / (tm["walk_kernel_ms"] / nq) / 1e6), flush=True)
        if TRACE and slots <= 4:
            E.query_batch("fora", queries[:slots], want_ppr=False)  # non-balanced: one launch holds every level
            out = np.zeros(4 * 4096, np.uint64)
            E.L.fora_debug_push_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
            lv = E.L.fora_debug_push_trace(E.h, out.ctypes.data, 4096)
            t = out[: 4 * lv].reshape(lv, 4).astype(np.int64)
            for i in range(lv - 1):
                dt = t[i + 1, 0] - t[i, 0]
                print("  L%3d nf=%8d E=%9d  level %8.1f us  phaseA %7.1f us  -> %6.2f G edges/s" % (i, t[i, 1], t[i, 2], dt / 1e3, (t[i, 3] - t[i, 0]) / 1e3, t[i, 2] / max(dt, 1)))
    finally:
        E.close()
