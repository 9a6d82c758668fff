"""Evaluate the BASELINE.json configurations other than the headline one and write profiles/r1_configs.json.

  config 1  webstanford-shape, plain FORA eps=0.5, 20 queries; the unmodified reference (oracle/_ref) on one core beside it
  config 3  Pokec-shape: build --opt index (sharded by source range when several GPUs are visible), then --with_idx --opt queries
  config 4  LiveJournal-shape top-k (k=500): fora (--opt and with bounds) / fwdpush / montecarlo / bippr against the
            power-iteration ground truth (gen-exact-topk semantics), precision as algo.h:524-572
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fora_b200 as fb

SHAPES = {"webstanford": (281904, 2312497), "pokec": (1632803, 30622564), "lj": (4847571, 68993773)}
EPS = 0.5
out = {}


def graph(shape, with_in=False):
    n, m = SHAPES[shape]
    src, dst = fb.synth_edges(n, m, 42)
    E = fb.Engine(0, seed=2026, slots=16)
    E.build_graph_from_edges(n, m, src, dst, with_in=with_in)
    return n, m, src, dst, E


def precision(nodes, exact_nodes, exact_vals, k):
    ex = set(int(v) for v, p in zip(exact_nodes[:k], exact_vals[:k]) if p > 0)
    return len(set(int(v) for v in nodes[:k]) & ex) / max(len(ex), 1)


which = sys.argv[1:] or ["1", "3", "4"]
if "1" in which:
    n, m, src, dst, E = graph("webstanford")
    q = np.random.default_rng(43).integers(0, n, 1000).astype(np.int32)[:20]
    E.configure("fora", EPS)
    E.query_batch("fora", q, want_ppr=False)
    t = time.perf_counter(); ppr, st, tm = E.query_batch("fora", q); dt = time.perf_counter() - t
    res = {"gpu_queries_per_s": 20 / (tm["total_ms"] / 1e3), "gpu_wall_queries_per_s_with_ppr_copy": 20 / dt, "push_ms": tm["push_ms"], "walk_ms": tm["walk_ms"]}
    import helpers
    if helpers.have_reference():
        op, oc, _, _ = fb.csr_from_edges(n, src, dst, with_in=False)
        class G: pass
        g = G(); g.n, g.m_decl, g.out_ptr, g.out_col = n, m, op, oc
        g.in_ptr, g.in_col = np.zeros(n + 1, np.int64), np.zeros(1, np.int32)
        R = helpers.Reference(g, epsilon=EPS)
        R.setting("fora"); R.init_query_state()
        for s in q: R.query("fora", int(s))
        res["reference_cpu_1core_queries_per_s"] = 20 / R.timer(3)
        res["reference_push_pct"], res["reference_walk_pct"] = 100 * R.timer(5) / R.timer(3), 100 * R.timer(6) / R.timer(3)
        # accuracy of both against GPU power iteration on 3 queries
        accs = []
        for i in range(3):
            exact = E.power_iteration(int(q[i]), 100)
            big = exact >= 1.0 / n
            R.query("fora", int(q[i]))
            accs.append((float((np.abs(ppr[i][big] - exact[big]) / exact[big]).max()), float((np.abs(R.ppr()[big] - exact[big]) / exact[big]).max())))
        res["max_rel_err_gpu_vs_reference"] = accs
    out["config1_webstanford"] = res
    print(json.dumps(res), flush=True)
    E.close()

if "3" in which:
    import torch
    n, m, src, dst, E = graph("pokec")
    ngpu = torch.cuda.device_count()
    rmax, omega = E.configure("fora", EPS, opt=1)
    off, cnt, total = E.index_info()
    engines = [E]
    for d in range(1, ngpu):
        Ed = fb.Engine(d, seed=2026, slots=16); Ed.build_graph_from_edges(n, m, src, dst, with_in=False); Ed.configure("fora", EPS, opt=1); engines.append(Ed)
    from fora_b200 import shard
    cuts = shard.balanced_source_ranges(off, cnt, ngpu)
    import threading
    parts = [None] * ngpu
    def work(d): parts[d] = engines[d].index_build(off, cnt, cuts[d], cuts[d + 1])
    th = [threading.Thread(target=work, args=(d,)) for d in range(ngpu)]  # warm-up (first-call allocations)
    [x.start() for x in th]; [x.join() for x in th]
    t = time.perf_counter()
    th = [threading.Thread(target=work, args=(d,)) for d in range(ngpu)]
    [x.start() for x in th]; [x.join() for x in th]
    t_build = time.perf_counter() - t
    dest = np.concatenate(parts)
    assert len(dest) == total
    E.index_upload(off, cnt, dest)
    E.configure("fora", EPS, opt=1, with_idx=1)
    q = np.random.default_rng(43).integers(0, n, 1000).astype(np.int32)[:200]
    E.query_batch("fora", q[:32], want_ppr=False)
    _, st, tm = E.query_batch("fora", q, want_ppr=False)
    hit = sum(s["n_idx_hits"] for s in st) / max(1, sum(s["n_walks"] for s in st))
    E.configure("fora", EPS, opt=1, with_idx=0)
    _, st2, tm2 = E.query_batch("fora", q, want_ppr=False)
    ppr, _, _ = (lambda r: r)(E.query_batch("fora", q[:3]))
    E.configure("fora", EPS, opt=1, with_idx=1)
    ppr_i, _, _ = E.query_batch("fora", q[:3])
    acc = []
    for i in range(3):
        exact = E.power_iteration(int(q[i]), 100); big = exact >= 1.0 / n
        acc.append(float((np.abs(ppr_i[i][big] - exact[big]) / exact[big]).max()))
    res = {"gpus": ngpu, "index_entries": int(total), "index_build_s": t_build, "index_walks_per_s": total / t_build,
           "with_idx_queries_per_s": 200 / (tm["total_ms"] / 1e3), "idx_hit_ratio": hit, "without_idx_queries_per_s": 200 / (tm2["total_ms"] / 1e3),
           "max_rel_err_with_idx": acc}
    out["config3_pokec"] = res
    print(json.dumps(res), flush=True)
    for e in engines: e.close()

if "4" in which:
    n, m, src, dst, E = graph("lj", with_in=True)
    k = 500
    q = np.random.default_rng(43).integers(0, n, 1000).astype(np.int32)[:16]
    exact = []
    t = time.perf_counter()
    E.configure("fora", EPS)
    for s in q:
        p = E.power_iteration(int(s), 100)
        exact.append(E.topk_of(p, k))
    res = {"k": k, "queries": len(q), "gen_exact_topk_s_per_query": (time.perf_counter() - t) / len(q), "algos": {}}
    for name, algo, kw, nq in (("fora --opt", "fora", dict(opt=1), 16), ("fora (bounds)", "fora", dict(opt=0), 8), ("fwdpush", "fwdpush", {}, 16),
                               ("montecarlo", "montecarlo", {}, 8), ("bippr", "bippr", {}, 2)):
        E.configure(algo, EPS, k=k, **kw)
        E.topk_batch(algo, q[:min(nq, 2)], k)  # warm-up: first-call allocations are not query time
        t = time.perf_counter()
        nodes, vals, iters, st, tm = E.topk_batch(algo, q[:nq], k)
        dt = time.perf_counter() - t
        pr = [precision(nodes[i], exact[i][0], exact[i][1], k) for i in range(nq)]
        res["algos"][name] = {"queries": nq, "avg_precision": float(np.mean(pr)), "s_per_query": dt / nq, "avg_iters": float(np.mean(iters))}
        print(name, res["algos"][name], flush=True)
    out["config4_lj_topk"] = res
    E.close()

path = os.path.join(ROOT, "gpurun_out", "r1_configs.json")
os.makedirs(os.path.dirname(path), exist_ok=True)
prev = {}
if os.path.exists(path):
    prev = json.load(open(path))
prev.update(out)
json.dump(prev, open(path, "w"), indent=1)
print("wrote", path)
