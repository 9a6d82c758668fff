"""compute-sanitizer workload: the opt-in paths on a small graph (development script).
usage: compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fora_b200 as fb
from helpers import Graph
g = Graph.synth(30000, 400000, seed=17)
E = fb.Engine(0, seed=3, slots=6)
E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col)
rmax, _ = E.configure("fora", 0.5, opt=1, balanced=1)
srcs = np.array([0, 11, int(np.argmax(g.deg)), int(np.flatnonzero(g.deg == 0)[0]), 5, 123, 77, 20000], np.int32)
ppr, st, tm = E.query_batch("fora", srcs)
print("queries ok", [round(float(p.sum()), 9) for p in ppr][:3], tm["kernel_launches"])
E.set_shared_walks(True)
ppr, st, tm = E.query_batch("fora", srcs)
print("shared ok", [round(float(p.sum()), 9) for p in ppr][:3], st[0]["n_idx_hits"] == st[0]["n_walks"])
E.close()
if os.environ.get("SANITIZE_ALL"):
    g = Graph.synth(6000, 70000, seed=5)
    E = fb.Engine(0, seed=3, slots=4)
    E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col, g.in_ptr, g.in_col)
    srcs = np.array([1, 7, int(np.argmax(g.deg)), 900, 2500], np.int32)
    for algo, kw in (("fora", dict(opt=1)), ("fora", dict(opt=0)), ("fwdpush", {}), ("montecarlo", {}), ("bippr", {})):
        E.configure(algo, 0.5, k=50, **kw)
        out = E.topk_batch(algo, srcs[:3], 50)
        print("topk", algo, kw, "ok")
    E.configure("bippr", 0.5)
    E.query_batch("bippr", srcs[:2])
    E.configure("montecarlo", 0.5)
    E.query_batch("montecarlo", srcs[:2])
    E.configure("fora", 0.5, opt=1, with_idx=1)
    off, cnt, total = E.index_info()
    dest = E.index_build(off, cnt)
    E.index_upload(off, cnt, dest)
    E.query_batch("fora", srcs)
    E.power_iteration(7, 20)
    ids, vals, offs, st, tm = E.query_batch_sparse("fora", srcs, 1.0 / g.n, g.n)
    print("all ok", total, int(offs[-1]))
    E.close()
