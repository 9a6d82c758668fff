"""One-process sweep of runtime variants of the hot path on the LJ-shape workload (development aid).

Every variant is an environment setting read by libfora_b200.so when a context is created, a graph uploaded or a
wave launched, so one GPU call can compare many of them: a fresh Engine per variant, 32 warm-up queries, then
`NQ` timed queries (FORA eps=0.5 --balanced --opt, results left on the device).  Greedy: a later group starts from
the best setting of the earlier groups.

usage: python scripts/variants.py [slots] [nq]   -> lines on stdout and gpurun_out/variants.jsonl
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fora_b200 as fb  # noqa: E402

if os.environ.get("FORA_VARIANT_LIB"):  # a build with other compile-time constants (fora_b200/variants/lib_<name>.so)
    fb.LIB_PATH = os.path.join(ROOT, "fora_b200", "variants", "lib_%s.so" % os.environ["FORA_VARIANT_LIB"])
    print("library:", fb.LIB_PATH, flush=True)

SLOTS = int(sys.argv[1]) if len(sys.argv) > 1 else 32
NQ = int(sys.argv[2]) if len(sys.argv) > 2 else 96
KNOBS = ("FORA_L2_FETCH", "FORA_PUSH_PACK", "FORA_RELABEL_KEY", "FORA_WALK_HOT_MB", "FORA_L2_HINTS", "FORA_NO_WALK_PIN",
         "FORA_COST_WALK", "FORA_COST_EDGE", "FORA_COST_VERTEX", "FORA_TILE_MAX", "FORA_WALK_GRID", "FORA_WALK_V", "FORA_WALK_HOT_KEEP",
         "FORA_PUSH_V", "FORA_PUSH_DYN", "FORA_DEBUG_NO_RED", "FORA_WALK_PIN_PPR", "FORA_PUSH_LOG", "FORA_PUSH_LOG_CAP")

n, m = 4847571, 68993773
t0 = time.time()
src, dst = fb.synth_edges(n, m, 42)
op, oc, _, _ = fb.csr_from_edges(n, src, dst, with_in=False)
del src, dst
queries = np.random.default_rng(43).integers(0, n, 1000).astype(np.int32)
print("graph ready in %.1f s" % (time.time() - t0), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "variants.jsonl"), "a")


def measure(name, env, slots=SLOTS):
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    E = fb.Engine(0, seed=2026, slots=slots)
    try:
        E.upload_graph(n, m, op, oc)
        E.configure("fora", 0.5, opt=1, balanced=1)
        E.query_batch("fora", queries[:slots], want_ppr=False)
        _, stats, tm = E.query_batch("fora", queries[slots:slots + NQ], want_ppr=False)
    finally:
        E.close()
    r = {
        "name": name, "env": env, "slots": slots, "nq": NQ,
        "qps": NQ / (tm["total_ms"] * 1e-3),
        "push_kernel_ms_per_q": tm["push_kernel_ms"] / NQ, "walk_kernel_ms_per_q": tm["walk_kernel_ms"] / NQ,
        "push_ms_per_q": tm["push_ms"] / NQ, "walk_ms_per_q": tm["walk_ms"] / NQ,
        "edges_per_q": sum(s["edges_pushed"] for s in stats) / NQ, "walks_per_q": sum(s["n_walks"] for s in stats) / NQ,
        "hops_per_q": sum(s["walk_hops"] for s in stats) / NQ,
    }
    r["G_edges_per_s"] = r["edges_per_q"] / r["push_kernel_ms_per_q"] / 1e6
    r["G_hops_per_s"] = r["hops_per_q"] / r["walk_kernel_ms_per_q"] / 1e6
    print("%-34s %7.1f q/s | push %.3f ms/q (%.1f G edges/s) walk %.3f ms/q (%.1f G hops/s) | E %.1fM W %.1fM" % (
        name, r["qps"], r["push_kernel_ms_per_q"], r["G_edges_per_s"], r["walk_kernel_ms_per_q"], r["G_hops_per_s"],
        r["edges_per_q"] / 1e6, r["walks_per_q"] / 1e6), flush=True)
    LOG.write(json.dumps(r) + "\n")
    LOG.flush()
    return r


def group(title, base, candidates):
    """measure base+candidate for every candidate, return the best env (or base if nothing beats it)"""
    print("--", title, flush=True)
    best_env, best = base, measure("base " + json.dumps(base), base)["qps"]
    for name, delta in candidates:
        env = dict(base)
        env.update(delta)
        q = measure(name, env)["qps"]
        if q > best * 1.005:
            best, best_env = q, env
    print("   best:", best_env, "%.1f q/s" % best, flush=True)
    return best_env


PLAN = sys.argv[3] if len(sys.argv) > 3 else "b"
env = {}
if PLAN == "a":  # first sweep of the session, kept for the record (FORA_L2_FETCH no longer exists in the engine: no effect)
    OFF = {"FORA_PUSH_PACK": 0, "FORA_RELABEL_KEY": 0, "FORA_TILE_MAX": 16384, "FORA_WALK_GRID": 8}
    env = group("L2 fill granularity", OFF, [("fetch32", dict(OFF, FORA_L2_FETCH=32)), ("fetch128", dict(OFF, FORA_L2_FETCH=128)),
                                              ("fetch64", dict(OFF, FORA_L2_FETCH=64))])
    env = group("push: out-degree packed into the column ids", env, [("pack", {"FORA_PUSH_PACK": 1})])
    env = group("relabel key", env, [("key=indeg/outdeg", {"FORA_RELABEL_KEY": 1})])
    env = group("walk: static hot prefix of the column array", env,
                [("hot %d MB" % mb, {"FORA_WALK_HOT_MB": mb}) for mb in (8, 16, 32, 48, 64)])
    env = group("misc", env, [("no push L2 hints", {"FORA_L2_HINTS": 0}), ("no row-offset pin", {"FORA_NO_WALK_PIN": 1}),
                              ("walk grid x16", {"FORA_WALK_GRID": 16}), ("tile 8K", {"FORA_TILE_MAX": 8192}),
                              ("tile 32K", {"FORA_TILE_MAX": 32768})])
    # re-balance push against walks around the calibrated cost model (engine.cu DEFAULT_COST_*)
    env = group("cost model", env, [("walk cost x0.8", {"FORA_COST_WALK": 6.5e-11 * 0.8}), ("walk cost x1.25", {"FORA_COST_WALK": 6.5e-11 * 1.25}),
                                    ("edge cost x0.7", {"FORA_COST_EDGE": 3.5e-11 * 0.7}), ("edge cost x1.4", {"FORA_COST_EDGE": 3.5e-11 * 1.4})])
elif PLAN == "b":
    env = group("tiles / walk grid", env, [("tile 64K", {"FORA_TILE_MAX": 65536}), ("tile 128K", {"FORA_TILE_MAX": 131072}),
                                           ("walk grid x8", {"FORA_WALK_GRID": 8}), ("walk grid x32", {"FORA_WALK_GRID": 32})])
    env = group("extra", env, [(k, json.loads(v)) for k, v in (a.split("=", 1) for a in sys.argv[4:])])
elif PLAN == "c":
    env = group("walk loop", env, [("walk v2 (warp-converged)", {"FORA_WALK_V": 2})])
    env = group("walk: cold column slots evict_first", env, [("cold beyond %d MB" % mb, {"FORA_WALK_HOT_MB": mb}) for mb in (4, 8, 16, 32, 64, 128)])
    env = group("extra", env, [(k, json.loads(v)) for k, v in (a.split("=", 1) for a in sys.argv[4:])])
elif PLAN == "m":  # one measurement of the defaults (used to compare variant libraries)
    measure("defaults", env)
    sys.exit(0)
elif PLAN == "x":  # extras only
    env = group("extra", env, [(k, json.loads(v)) for k, v in (a.split("=", 1) for a in sys.argv[4:])])
elif PLAN == "d":
    env = group("push: dynamic guided tiles", env, [("dyn 32K", {"FORA_PUSH_DYN": 1}), ("dyn 16K", {"FORA_PUSH_DYN": 1, "FORA_TILE_MAX": 16384}),
                                                    ("dyn 64K", {"FORA_PUSH_DYN": 1, "FORA_TILE_MAX": 65536}),
                                                    ("dyn 8K", {"FORA_PUSH_DYN": 1, "FORA_TILE_MAX": 8192})])
    env = group("extra", env, [(k, json.loads(v)) for k, v in (a.split("=", 1) for a in sys.argv[4:])])
print("-- slots", flush=True)
for s in ((16, 24, 40, 48, 56, 64) if PLAN in "ab" else (56,)):
    measure("slots %d" % s, env, slots=s)
print("FINAL", json.dumps(env), flush=True)

# the winning configuration must give the same answers: push to 1e-12 (atomic order only), PPR mass 1, walk counts equal
def answers(e):
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in e.items()})
    E = fb.Engine(0, seed=2026, slots=4)
    try:
        E.upload_graph(n, m, op, oc)
        rmax, _ = E.configure("fora", 0.5, opt=1, balanced=1)
        res, rsd, rsum, _ = E.push_only(int(queries[0]), rmax * 4)
        ppr, stats, _ = E.query_batch("fora", queries[:4])
    finally:
        E.close()
    return res, rsd, rsum, ppr, [s["n_walks"] for s in stats]


a, b = answers({}), answers(env)
print("check: push |d reserve| %.2e |d residue| %.2e |d rsum| %.2e; ppr sums %s; max |d ppr| %.2e; walks %s vs %s" % (
    np.abs(a[0] - b[0]).max(), np.abs(a[1] - b[1]).max(), abs(a[2] - b[2]), np.round(b[3].sum(1), 12).tolist(),
    np.abs(a[3] - b[3]).max(), a[4], b[4]), flush=True)
