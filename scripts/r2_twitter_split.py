"""BASELINE config 5: ONE whole-graph SSPPR query with the walks split over 1/2/4/8 GPUs of a box, in one process through the
C-ABI group API (fora_group_*: push on GPU 0, ncclBroadcast of the compacted residue list, walks by chunk range, ncclAllReduce).

usage: python scripts/r2_twitter_split.py [shape] [group sizes ...]     (shape: twitter | lj | small)
One JSON line per (group size, balanced) with the phase breakdown; results are compared across group sizes and with exact PPR.
"""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fora_b200 as fb  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "lj"
sizes = [int(a) for a in sys.argv[2:]] or [1, 2]
n, m = {"lj": (4847571, 68993773), "small": (200000, 3000000), "twitter": (41652230, 1468365182)}[shape]
t0 = time.time()
src, dst = fb.synth_edges(n, m, 42)
print("synth %.1fs" % (time.time() - t0), flush=True)
s = None
ref = {}
for g in sizes:
    Gg = fb.Group(g, seed=99)  # a fresh group per size: its NCCL communicator spans exactly g GPUs
    t0 = time.time()
    th = [threading.Thread(target=lambda E=E: E.build_graph_from_edges(n, m, src, dst, with_in=False)) for E in Gg.engines]  # K0 on every GPU
    [t.start() for t in th]
    [t.join() for t in th]
    print("graph built on %d GPU(s) in %.1fs" % (g, time.time() - t0), flush=True)
    if s is None:
        deg = np.diff(Gg.engines[0].download_csr(with_in=False)[0])
        s = int(np.flatnonzero(deg > 5)[4321])
        del deg
    for bal in (0, 1):
        rmax, omega = Gg.configure("fora", 0.5, opt=1, balanced=bal)
        best = None
        for rep in range(4):  # timed without shipping the vector to the host (the reference discards it, query.h:1471-1476)
            _, st, tm = Gg.query_split(s, query_id=0, want_ppr=False)
            if rep and (best is None or tm["total_ms"] < best[1]["total_ms"]):
                best = (st, tm)
        st, tm = best
        ppr = Gg.query_split(s, query_id=0)[0]
        out = {"shape": shape, "n": n, "m": m, "gpus": g, "balanced": bal, "source": s, "sum": float(ppr.sum())}
        out.update({k: (round(v, 3) if isinstance(v, float) else v) for k, v in tm.items()})
        out.update({"walks": st["n_walks"], "hops": st["walk_hops"], "edges_pushed": st["edges_pushed"], "push_rounds": st["push_rounds"], "rsum": st["rsum"]})
        if bal == 0:
            if "base" not in ref:
                ref["base"] = ppr
            out["max_abs_diff_vs_first_group_size"] = float(np.abs(ppr - ref["base"]).max())
        if g == sizes[-1]:
            exact = Gg.engines[0].power_iteration(s, 100)
            big = exact >= 1.0 / n
            out["max_rel_err_vs_exact_on_pi_ge_1_over_n"] = float((np.abs(ppr[big] - exact[big]) / exact[big]).max())
        print(json.dumps(out), flush=True)
    Gg.close()
