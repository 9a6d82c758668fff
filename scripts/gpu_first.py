"""First end-to-end GPU check (development script; the formal checks live in tests/)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fora_b200 as fb
from helpers import Graph, Oracle

def relerr(a, b):
    d = np.abs(a - b); s = np.maximum(np.abs(a), np.abs(b)); m = s > 0
    return float((d[m] / s[m]).max()) if m.any() else 0.0

def small():
    g = Graph.synth(20000, 200000, seed=3)
    E = fb.Engine(0, seed=7, slots=4)
    E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col, g.in_ptr, g.in_col)
    op, oc, ip_, ic = E.download_csr()
    print("csr roundtrip", np.array_equal(op, g.out_ptr), np.array_equal(oc, g.out_col), np.array_equal(ip_, g.in_ptr), np.array_equal(ic, g.in_col))
    rmax, omega = E.configure("fora", 0.5)
    O = Oracle(g); O.set_params(0.5, rmax, omega); O.init_state(-1.0, 0)
    for s in [0, 11, int(np.argmax(g.deg)), int(np.flatnonzero(g.deg == 0)[0])]:
        res, rsd, rsum, st = E.push_only(s, rmax)
        O.reset_counters()
        r2 = O.push_sync(s, rmax, 1, 0); a, b = O.fwd(); c = O.counters()
        print("push s=%d deg=%d rsum %.12g/%.12g reserve relerr %.2e residue relerr %.2e mass %.15f edges %d/%d verts %d/%d levels %d/%d" % (
            s, g.deg[s], rsum, r2, relerr(res, a), relerr(rsd, b), res.sum() + rsd.sum(), st["edges_pushed"], c["edges_pushed"],
            st["vertices_pushed"], c["vertices_pushed"], st["push_levels"], c["push_levels"]))
    # resumable rounds
    s = 11
    E.push_begin(s); O.push_topk_begin(s)
    for rm in [rmax * 8, rmax * 4, rmax * 2, rmax]:
        res, rsd, rsum, st = E.push_round(rm)
        r2 = O.push_sync(s, rm, 0, 1); a, b = O.fwd()
        print("round rmax=%.3g rsum %.12g/%.12g relerr %.2e %.2e" % (rm, rsum, r2, relerr(res, a), relerr(rsd, b)))
    # walks
    d, hops = E.random_walks(11, 1000000, 0)
    print("walk hops/walk %.4f (expect 4 minus dangling effects)" % (hops / 1e6))
    d2, hops2 = E.random_walks(11, 1000000, 1)
    print("no-zero-hop hops/walk %.4f (expect ~5)" % (hops2 / 1e6))
    # fora query vs power iteration
    for opt, bal in [(0, 0), (1, 0), (1, 1)]:
        rmax, omega = E.configure("fora", 0.5, opt=opt, balanced=bal)
        srcs = np.array([0, 11, 5, int(np.argmax(g.deg)), 123, 77], np.int32)
        ppr, stats, tm = E.query_batch("fora", srcs)
        for i, s in enumerate(srcs[:3]):
            exact = E.power_iteration(int(s), 100)
            ex2 = O.power_iteration(int(s), 100)
            m = exact >= 1.0 / g.n
            rel = np.abs(ppr[i][m] - exact[m]) / exact[m]
            print("opt=%d bal=%d s=%d sum(ppr)=%.6f max rel err on pi>=delta: %.3f (eps=0.5) walks=%d rsum=%.4f rounds=%d pi-vs-oracle %.2e" % (
                opt, bal, s, ppr[i].sum(), rel.max(), stats[i]["n_walks"], stats[i]["rsum"], stats[i]["push_rounds"], np.abs(exact - ex2).max()))
        print("timing", tm)
    E.close()

def big(n, m, nq, slots, opt=1, balanced=1):
    t0 = time.time()
    src, dst = fb.synth_edges(n, m, 42)
    t1 = time.time()
    op, oc, _, _ = fb.csr_from_edges(n, src, dst, with_in=False)
    t2 = time.time()
    print("synth %.1fs csr %.1fs maxdeg %d dangling %d" % (t1 - t0, t2 - t1, np.diff(op).max(), (np.diff(op) == 0).sum()))
    E = fb.Engine(0, seed=7, slots=slots)
    E.upload_graph(n, m, op, oc)
    rmax, omega = E.configure("fora", 0.5, opt=opt, balanced=balanced)
    print("rmax %.4g omega %.4g" % (rmax, omega))
    rng = np.random.default_rng(43)
    srcs = rng.integers(0, n, nq).astype(np.int32)
    for rep in range(2):
        t0 = time.time()
        _, stats, tm = E.query_batch("fora", srcs, want_ppr=False)
        dt = time.time() - t0
        W = sum(s["n_walks"] for s in stats); H = sum(s["walk_hops"] for s in stats); Ed = sum(s["edges_pushed"] for s in stats)
        V = sum(s["vertices_pushed"] for s in stats)
        print("slots=%d nq=%d wall %.3fs -> %.1f q/s | push %.1f ms walk %.1f ms | walks %.3g hops %.3g (%.2f G hops/s) edges %.3g (%.2f G edges/s) verts %.3g rounds %.1f levels %.1f rsum %.3f launches %d" % (
            slots, nq, dt, nq / dt, tm["push_ms"], tm["walk_ms"], W, H, H / tm["walk_ms"] / 1e6, Ed, Ed / max(tm["push_ms"], 1e-9) / 1e6, V,
            np.mean([s["push_rounds"] for s in stats]), np.mean([s["push_levels"] for s in stats]), np.mean([s["rsum"] for s in stats]), tm["kernel_launches"]))
    E.close()

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "small"
    if what == "small":
        small()
    elif what == "lj":
        for slots in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "4").split(",")]:
            big(4847571, 68993773, int(sys.argv[3]) if len(sys.argv) > 3 else 16, slots)
    elif what == "ws":
        big(281904, 2312497, 32, 8, opt=0, balanced=0)
