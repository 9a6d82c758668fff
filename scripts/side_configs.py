"""Warm, device-timed measurements of the non-headline workloads (development aid; compare libraries with FORA_VARIANT_LIB).
webstanford-shape plain FORA, LJ-shape top-k (--opt and with bounds), Pokec-shape --with_idx, index build."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fora_b200 as fb
if os.environ.get("FORA_VARIANT_LIB"):
    fb.LIB_PATH = os.path.join(ROOT, "fora_b200", "variants", "lib_%s.so" % os.environ["FORA_VARIANT_LIB"])
SH = {"webstanford": (281904, 2312497), "pokec": (1632803, 30622564), "lj": (4847571, 68993773)}
which = sys.argv[1:] or ["ws", "topk", "pokec"]

def eng(shape, slots=16, with_in=False):
    n, m = SH[shape]
    src, dst = fb.synth_edges(n, m, 42)
    E = fb.Engine(0, seed=2026, slots=slots)
    E.build_graph_from_edges(n, m, src, dst, with_in=with_in)
    return n, E

def show(name, nq, tm):
    print("%-34s %8.3f ms/query (push %.3f walk %.3f topk %.3f ms/query, %d launches)" % (
        name, tm["total_ms"] / nq, tm["push_ms"] / nq, tm["walk_ms"] / nq, tm.get("topk_ms", 0) / nq, tm["kernel_launches"]), flush=True)

if "ws" in which:
    n, E = eng("webstanford")
    q = np.random.default_rng(43).integers(0, n, 1000).astype(np.int32)
    for name, kw in (("webstanford fora", dict()), ("webstanford fora --opt --balanced", dict(opt=1, balanced=1))):
        E.configure("fora", 0.5, **kw)
        E.query_batch("fora", q[:32], want_ppr=False)
        _, st, tm = E.query_batch("fora", q[32:96], want_ppr=False)
        show(name, 64, tm)
    E.close()
if "topk" in which:
    n, E = eng("lj")
    q = np.random.default_rng(43).integers(0, n, 1000).astype(np.int32)
    for name, algo, kw, nq in (("lj topk fora --opt", "fora", dict(opt=1), 32), ("lj topk fora (bounds)", "fora", dict(opt=0), 16),
                               ("lj topk fwdpush", "fwdpush", {}, 16)):
        E.configure(algo, 0.5, k=500, **kw)
        E.topk_batch(algo, q[:nq], 500)
        *_, st, tm = E.topk_batch(algo, q[nq:2 * nq], 500)
        show(name, nq, tm)
    E.close()
if "mc" in which:  # Monte-Carlo baseline (query.h:16-43) and BiPPR's walk phase run through the chunked walk kernel
    n, E = eng("lj")
    q = np.random.default_rng(43).integers(0, n, 1000).astype(np.int32)
    rmax, omega = E.configure("montecarlo", 0.5)
    E.query_batch("montecarlo", q[:1], want_ppr=False)
    _, st, tm = E.query_batch("montecarlo", q[1:5], want_ppr=False)
    hops = sum(s["walk_hops"] for s in st)
    print("lj montecarlo %.1f ms/query, walk kernels %.1f ms/query: %.1f G hops/s (%.3g walks per query)" % (
        tm["total_ms"] / 4, tm["walk_kernel_ms"] / 4, hops / tm["walk_kernel_ms"] / 1e6, st[0]["n_walks"]), flush=True)
    E.close()
if "pokec" in which:
    n, E = eng("pokec")
    q = np.random.default_rng(43).integers(0, n, 1000).astype(np.int32)
    E.configure("fora", 0.5, opt=1, with_idx=1)
    off, cnt, total = E.index_info()
    E.index_build(off, cnt)
    t = time.perf_counter(); dest = E.index_build(off, cnt); dt = time.perf_counter() - t
    walks, hops, kms = E.index_build_stat()
    # index build bytes = 20*hops + 4*walks
    print("pokec index build %.1f ms wall (%d entries, incl. D2H of the index into pageable memory) | walk kernels %.2f ms: %.1f G hops/s, %.0f GB/s algorithmic (20*hops+4*walks)" % (
        dt * 1e3, total, kms, hops / kms / 1e6, (20.0 * hops + 4.0 * walks) / kms / 1e6), flush=True)
    E.index_upload(off, cnt, dest)
    E.query_batch("fora", q[:32], want_ppr=False)
    _, st, tm = E.query_batch("fora", q[32:160], want_ppr=False)
    show("pokec fora --opt --with_idx", 128, tm)
    E.configure("fora", 0.5, opt=1, with_idx=0)
    E.query_batch("fora", q[:32], want_ppr=False)
    _, st, tm = E.query_batch("fora", q[32:160], want_ppr=False)
    show("pokec fora --opt (no index)", 128, tm)
    E.close()
