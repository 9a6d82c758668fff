"""GPU tests of the round-2 additions to the C ABI and the kernels behind them: compacted output, global query index,
CSR validation on upload, the backward push's touched list at small r_max, bulk walks through the chunked walk kernel
(index build statistics), and the second-generation push (sub-waves + tails) against the first-generation kernel."""
import os
import subprocess
import sys

import numpy as np
import pytest

import fora_b200 as fb
from helpers import Graph, Oracle

pytestmark = pytest.mark.gpu
EPS = 0.5
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def relerr(a, b):
    d = np.abs(a - b)
    s = np.maximum(np.abs(a), np.abs(b))
    m = s > 0
    return float((d[m] / s[m]).max()) if m.any() else 0.0


@pytest.fixture(scope="module")
def g():
    return Graph.synth(20000, 200000, seed=3)


def test_sparse_output_matches_dense(g):
    # fora_query_batch_sparse: the (id, value) pairs >= threshold are exactly the dense vector's (the reference discards
    # the vector, query.h:1471-1476; FORA's guarantee covers pi >= 1/n)
    E = fb.Engine(0, seed=5, slots=3)
    E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col)
    E.configure("fora", EPS, opt=1, balanced=1)
    srcs = np.array([0, 11, int(np.argmax(g.deg)), int(np.flatnonzero(g.deg == 0)[0]), 5, 123, 77], np.int32)
    ppr, stats, _ = E.query_batch("fora", srcs)
    thr = 1.0 / g.n
    ids, vals, off, st2, _ = E.query_batch_sparse("fora", srcs, thr, g.n)
    assert off[0] == 0 and len(off) == len(srcs) + 1
    for i in range(len(srcs)):
        sel, v = ids[int(off[i]):int(off[i + 1])], vals[int(off[i]):int(off[i + 1])]
        assert len(np.unique(sel)) == len(sel)
        must = np.flatnonzero(ppr[i] >= thr * (1 + 1e-9))  # entries within rounding of the threshold may fall either side
        may = np.flatnonzero(ppr[i] >= thr * (1 - 1e-9))
        assert np.isin(must, sel).all() and np.isin(sel, may).all()
        assert np.allclose(v, ppr[i][sel], rtol=1e-9, atol=0)
        assert st2[i]["n_walks"] == stats[i]["n_walks"] and st2[i]["walk_hops"] == stats[i]["walk_hops"]
    # several waves (7 queries, 3 slots) landed in one packed buffer; capacities are enforced
    with pytest.raises(fb.ForaError):
        E.query_batch_sparse("fora", srcs, thr, 4)
    with pytest.raises(fb.ForaError):
        E.query_batch_sparse("fora", srcs, thr, g.n, ids=np.empty(10, np.int32), vals=np.empty(10))
    with pytest.raises(fb.ForaError):
        E.query_batch_sparse("fora", srcs, 0.0, g.n)
    E.close()


def test_query_base_keys_philox_by_global_index(g):
    # a query list cut into shards / successive calls gives the same walks as one call, and the same
    # source at another list position gets another stream
    srcs = np.array([0, 11, 5, 123, 77, 9, 4000], np.int32)
    E = fb.Engine(0, seed=21, slots=4)
    E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col)
    E.configure("fora", EPS, opt=1)
    _, whole, _ = E.query_batch("fora", srcs, want_ppr=False)
    E.set_query_base(3)
    _, part, _ = E.query_batch("fora", srcs[3:], want_ppr=False)
    assert [s["walk_hops"] for s in part] == [s["walk_hops"] for s in whole[3:]]
    E.set_query_base(0)
    _, again, _ = E.query_batch("fora", srcs[3:], want_ppr=False)
    assert [s["walk_hops"] for s in again] != [s["walk_hops"] for s in whole[3:]]
    assert [s["n_walks"] for s in again] == [s["n_walks"] for s in whole[3:]]
    E.close()


def test_upload_rejects_malformed_csr(g):
    E = fb.Engine(0)
    bad_col = g.out_col.copy()
    bad_col[17] = g.n  # id outside [0, n)
    with pytest.raises(fb.ForaError, match="column id"):
        E.upload_graph(g.n, g.m_decl, g.out_ptr, bad_col)
    bad_ptr = g.out_ptr.copy()
    bad_ptr[100], bad_ptr[101] = bad_ptr[101] + 5, bad_ptr[100]  # not non-decreasing
    with pytest.raises(fb.ForaError):
        E.upload_graph(g.n, g.m_decl, bad_ptr, g.out_col)
    with pytest.raises(fb.ForaError):  # in-CSR with another edge count
        E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col, g.in_ptr[: g.n + 1] // 2, g.in_col)
    E.n, E.m_decl = g.n, g.m_decl
    E.configure("fora", EPS)
    with pytest.raises(fb.ForaError, match="no graph"):  # nothing half-initialised is left behind
        E.push_only(0, 1e-6)
    E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col)  # and the context still works
    rmax, _ = E.configure("fora", EPS)
    res, rsd, rsum, _ = E.push_only(3, rmax)
    assert abs(res.sum() + rsd.sum() - 1) < 1e-12
    E.close()


def test_reverse_push_small_rmax_lists_each_vertex_once(g):
    # with a small r_max a vertex zeroed in phase A and hit again in phase B re-entered the touched list every
    # level and overflowed it; every vertex is now listed once per target.  Values against the synchronous oracle.
    E = fb.Engine(0)
    E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col, g.in_ptr, g.in_col)
    E.configure("bippr", EPS)
    O = Oracle(g)
    hub_in = int(np.argmax(np.diff(g.in_ptr)))
    for rm in (1e-6, 1e-8):
        for t in (hub_in, 11):
            res, rsd = E.reverse_push(t, rm)
            O.reverse_push(t, rm, 1.0, 1)
            a, b = O.bwd()
            assert relerr(res, a) < 1e-9 and relerr(rsd, b) < 1e-9, (rm, t)
    # the scratch is clean afterwards: a BiPPR query on the same context is still right
    rmax, omega = E.configure("bippr", EPS)
    ppr, stats, _ = E.query_batch("bippr", np.array([11], np.int32))
    exact = O.power_iteration(11, 150)
    big = exact >= 1.0 / g.n
    assert np.median(np.abs(ppr[0][big] - exact[big]) / exact[big]) < 0.15
    E.close()


def test_index_build_stats_and_bulk_walk_distribution(g):
    # the index build, Monte-Carlo and the walk test hook run through the chunked walk kernel of the query path
    E = fb.Engine(0, seed=9)
    E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col)
    for opt in (0, 1):
        rmax, omega = E.configure("fora", EPS, opt=opt)
        off, cnt, total = E.index_info()
        dest = E.index_build(off, cnt)
        walks, hops, ms = E.index_build_stat()
        assert walks == total == len(dest) and ms > 0
        nd = cnt[g.deg > 0].sum()
        # E[hops] = (1-a)/a = 4 (+1 with the forced first hop) minus the steps spent jumping back from dangling vertices (3 % of the nodes)
        assert (4.4 if opt else 3.5) < hops / max(nd, 1) < (5.02 if opt else 4.02)
        # a dangling source returns itself (algo.h:127-129); its slice is all "itself"
        dang = int(np.flatnonzero((g.deg == 0) & (cnt > 0))[0]) if ((g.deg == 0) & (cnt > 0)).any() else None
        if dang is not None:
            assert (dest[int(off[dang]):int(off[dang] + cnt[dang])] == dang).all()
        # per-source destination distribution of the slice of one hub against the oracle's walks (chi-square)
        hub = int(np.argmax(cnt))
        sl = dest[int(off[hub]):int(off[hub] + cnt[hub])]
        N = 200000
        d2, _ = E.random_walks(hub, N, opt)
        O = Oracle(g, seed=4)
        b = np.bincount(O.walks(hub, N, opt), minlength=g.n).astype(np.float64)
        a = np.bincount(d2, minlength=g.n).astype(np.float64)
        m = (a + b) >= 20
        chi2 = (((a - b)[m] ** 2) / (a + b)[m]).sum()
        dof = m.sum() - 1
        assert chi2 < dof + 6 * np.sqrt(2 * dof), (chi2, dof)
        assert sl.min() >= 0 and sl.max() < g.n
        # three shards = the unsharded build (Philox keyed by source and walk index)
        cuts = [0, g.n // 5, g.n // 2, g.n]
        assert np.array_equal(np.concatenate([E.index_build(off, cnt, cuts[i], cuts[i + 1]) for i in range(3)]), dest)
    E.close()


def test_shared_walks_keep_the_guarantee_and_walk_once(g):
    # opt-in per-wave walk pool (fora_ctx_set_shared_walks): every walk of every query is a hit in the pool, each query still meets
    # FORA's (eps, 1/n) guarantee against exact PPR, the push state is untouched, and the pool is smaller than the sum of the walks
    E = fb.Engine(0, seed=5, slots=8)
    E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col)
    E.configure("fora", 0.5, opt=1, balanced=1)
    srcs = np.array([3, 77, int(np.argmax(g.deg)), 1500, 9, 4000, 123, 2500, 31, 600], np.int32)  # two waves, the second ragged
    base, st0, _ = E.query_batch("fora", srcs)
    E.set_shared_walks(True)
    ppr, st1, _ = E.query_batch("fora", srcs)
    again, _, _ = E.query_batch("fora", srcs)
    assert np.abs(ppr - again).max() < 1e-12  # the pool is keyed by the wave's first global query index: reproducible
    bad = total = 0
    for i, s in enumerate(srcs):
        assert st1[i]["n_walks"] > 0 or g.deg[srcs[i]] == 0
        assert st1[i]["n_idx_hits"] == st1[i]["n_walks"] and st1[i]["walk_hops"] == 0
        assert abs(ppr[i].sum() - 1.0) < 1e-9
        exact = E.power_iteration(int(s), 150)
        big = exact >= 1.0 / g.n
        rel = np.abs(ppr[i][big] - exact[big]) / exact[big]
        bad += int((rel > 0.5).sum())
        total += int(big.sum())
        if st1[i]["n_walks"] > 1000:
            assert np.abs(ppr[i] - base[i]).max() > 0  # other walks than the private ones
    assert bad <= max(1, total // g.n), (bad, total)
    E.set_shared_walks(False)
    back, _, _ = E.query_batch("fora", srcs)
    assert np.abs(back - base).max() < 1e-12
    E.close()


def test_push_generations_agree():
    # the sub-wave / tail kernels (push2.cuh) and the first-generation kernel produce the same push: identical work
    # counters, values equal up to the order of fp64 additions; several sub-wave sizes and tail thresholds
    code = r"""
import os, sys, json
import numpy as np
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import fora_b200 as fb
from helpers import Graph
g = Graph.synth(60000, 900000, seed=17)
E = fb.Engine(0, seed=3, slots=6)
E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col)
rmax, _ = E.configure("fora", 0.5, opt=1, balanced=1)
srcs = np.array([0, 11, int(np.argmax(g.deg)), int(np.flatnonzero(g.deg == 0)[0]), 5, 123, 77, 31000], np.int32)
out = {}
res, rsd, rsum, st = E.push_only(int(srcs[2]), rmax)
out["push"] = [res.tolist(), rsd.tolist(), rsum, st["edges_pushed"], st["vertices_pushed"], st["push_levels"]]
E.push_begin(int(srcs[1]))
rounds = []
for k in range(4):
    res, rsd, rsum, st = E.push_round(rmax * 8 / 2 ** k)
    rounds.append([float(res.sum()), float(rsd.sum()), rsum, st["edges_pushed"], st["push_levels"]])
out["rounds"] = rounds
_, stats, _ = E.query_batch("fora", srcs, want_ppr=False)
out["stats"] = [[s["edges_pushed"], s["vertices_pushed"], s["push_levels"], s["push_rounds"], s["n_walks"], s["rsum"], s["final_rmax"]] for s in stats]
print("RESULT" + json.dumps(out))
""" % (ROOT, ROOT)
    import json

    def run(env):
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=e, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        return json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT")][0][6:])

    base = run({"FORA_PUSH_V": "1"})
    for env in ({"FORA_PUSH_V": "2", "FORA_PUSH_SUB": "2"}, {"FORA_PUSH_V": "2", "FORA_PUSH_SUB": "3", "FORA_TAIL_NF": "64", "FORA_TAIL_E": "300"},
                {"FORA_PUSH_V": "2", "FORA_PUSH_SUB": "1", "FORA_TAIL_NF": "2048", "FORA_TAIL_E": "100000"},
                {"FORA_PUSH_V": "2", "FORA_PUSH_SUB": "8", "FORA_TAIL_NF": "0", "FORA_PUSH_PREFETCH": "0"},
                # lockstep kernel (push3.cuh): one group per level, one slot per group, hubs cut into small pieces, plain atomics
                # first-generation kernel with dense slot-levels (RED + scan of the residue vector): always / from 1 % of the vertices
                {"FORA_PUSH_V": "1", "FORA_PUSH_DENSE": "0"}, {"FORA_PUSH_V": "1", "FORA_PUSH_DENSE": "0.01", "FORA_PUSH_LOG": "0"},
                {"FORA_PUSH_V": "1", "FORA_PUSH_DENSE": "0.005", "FORA_PUSH_EL": "0"},  # ... without the edge lists (tiles with RED)
                # ... and with a lockstep phase B (grid barrier between groups of slots on the edge line)
                {"FORA_PUSH_V": "1", "FORA_PUSH_LOCKSTEP": "0.05"}, {"FORA_PUSH_V": "1", "FORA_PUSH_LOCKSTEP": "0.5", "FORA_PUSH_DENSE": "0.02"},
                # (push3.cuh) default thresholds; every slot-level dense (RED + scan); never dense with tiny groups; hubs cut into
                # small pieces in both modes; plain atomics and no credit log
                {"FORA_PUSH_V": "3"}, {"FORA_PUSH_V": "3", "FORA_P3_DENSE": "0"}, {"FORA_PUSH_V": "3", "FORA_P3_DENSE": "-1", "FORA_P3_BUDGET": "0.002"},
                {"FORA_PUSH_V": "3", "FORA_P3_DENSE": "0.01", "FORA_P3_BUDGET": "0.05", "FORA_P3_HUB": "40,16"},
                {"FORA_PUSH_V": "3", "FORA_P3_DENSE": "0.001", "FORA_P3_BUDGET": "0.0001", "FORA_P3_HUB": "1,1", "FORA_L2_HINTS": "0", "FORA_PUSH_LOG": "0"}):
        got = run(env)
        assert got["push"][3:] == base["push"][3:], env
        assert relerr(np.array(got["push"][0]), np.array(base["push"][0])) < 1e-9 and relerr(np.array(got["push"][1]), np.array(base["push"][1])) < 1e-9
        assert abs(got["push"][2] - base["push"][2]) < 1e-12
        for a, b in zip(got["rounds"], base["rounds"]):
            assert a[3:] == b[3:] and np.allclose(a[:3], b[:3], rtol=0, atol=1e-12), env
        for a, b in zip(got["stats"], base["stats"]):
            assert a[:5] == b[:5] and abs(a[5] - b[5]) < 1e-12 and a[6] == b[6], (env, a, b)
