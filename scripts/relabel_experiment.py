"""Does degree-ordered relabelling pay?  Run the unchanged engine on the LJ-shape graph as generated and on the same
graph with vertices renumbered by descending in-degree (development experiment)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fora_b200 as fb
n, m = 4847571, 68993773
src, dst = fb.synth_edges(n, m, 42)
q = np.random.default_rng(43).integers(0, n, 1000).astype(np.int32)[:192]
def run(tag, src, dst, q, slots=32):
    E = fb.Engine(0, seed=2026, slots=slots)
    E.build_graph_from_edges(n, m, src, dst, with_in=False)
    E.configure("fora", 0.5, opt=1, balanced=1)
    E.query_batch("fora", q[:64], want_ppr=False)
    _, st, tm = E.query_batch("fora", q, want_ppr=False)
    E.close()
    ed = sum(s["edges_pushed"] for s in st); hp = sum(s["walk_hops"] for s in st)
    print("%-28s %.1f q/s | push kernel %.2f ms/q (%.1f G edges/s) walk kernel %.2f ms/q (%.1f G hops/s)" % (
        tag, len(q) / (tm["total_ms"] / 1e3), tm["push_kernel_ms"] / len(q), ed / tm["push_kernel_ms"] / 1e6, tm["walk_kernel_ms"] / len(q), hp / tm["walk_kernel_ms"] / 1e6), flush=True)
run("original ids", src, dst, q)
for key in ("in", "out", "in+out"):
    indeg = np.bincount(dst, minlength=n); outdeg = np.bincount(src, minlength=n)
    score = {"in": indeg, "out": outdeg, "in+out": indeg + outdeg}[key]
    new2old = np.argsort(-score, kind="stable").astype(np.int32)
    old2new = np.empty(n, np.int32); old2new[new2old] = np.arange(n, dtype=np.int32)
    # keep file order (edge order) so every adjacency list keeps its internal order
    run("relabelled by %s-degree" % key, old2new[src], old2new[dst], old2new[q])
