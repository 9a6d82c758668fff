"""GPU parity tests: the CUDA path, called through the C ABI (libfora_b200.so via ctypes), against the
CPU oracle on the same seeded inputs, against the committed golden vectors produced by the
unmodified reference, and -- at full LiveJournal-shape size -- through size-independent properties.

Tolerances: CSR / index layout / walk counts bit-exact; push reserve/residue 1e-9 relative against
the schedule-matched oracle (north_star allows 1e-6; only the order of fp64 atomic additions
differs); Monte-Carlo estimates within FORA's (eps, delta) guarantee against power iteration.
"""
import os

import numpy as np
import pytest

import fora_b200 as fb
from helpers import GOLDEN_DIR, Graph, Oracle

pytestmark = pytest.mark.gpu
EPS = 0.5
PUSH_RTOL = 1e-9


def relerr(a, b):
    d = np.abs(a - b)
    s = np.maximum(np.abs(a), np.abs(b))
    m = s > 0
    return float((d[m] / s[m]).max()) if m.any() else 0.0


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN_DIR, "ref_small.npz"))


@pytest.fixture(scope="module")
def g():
    return Graph.synth(20000, 200000, seed=3)


@pytest.fixture(scope="module")
def eng(g):
    E = fb.Engine(0, seed=7, slots=4)
    E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col, g.in_ptr, g.in_col)
    yield E
    E.close()


def interesting_sources(g):
    return [0, 11, int(np.argmax(g.deg)), int(np.flatnonzero(g.deg == 0)[0]), int(np.flatnonzero(g.deg == 1)[0])]


# ------------------------------------------------------------------------------------ graph
def test_csr_roundtrip_bit_exact(gold):
    n = int(gold["n"])
    E = fb.Engine(0)
    E.upload_graph(n, int(gold["m_decl"]), gold["out_ptr"], gold["out_col"], gold["in_ptr"], gold["in_col"])
    op, oc, ip_, ic = E.download_csr()
    assert np.array_equal(op, gold["out_ptr"]) and np.array_equal(oc, gold["out_col"])
    assert np.array_equal(ip_, gold["in_ptr"]) and np.array_equal(ic, gold["in_col"])
    E.close()


def test_device_csr_build_bit_exact(gold, g):
    # K0: count -> scan -> stable fill on the device == the reference's push_back order (graph.h:152-160)
    n = int(gold["n"])
    E = fb.Engine(0)
    E.build_graph_from_edges(n, int(gold["m_decl"]), gold["src"], gold["dst"])
    op, oc, ip_, ic = E.download_csr()
    assert np.array_equal(op, gold["out_ptr"]) and np.array_equal(oc, gold["out_col"])
    assert np.array_equal(ip_, gold["in_ptr"]) and np.array_equal(ic, gold["in_col"])
    g3 = Graph.synth(30000, 400000, seed=9, self_loops=50)
    E.build_graph_from_edges(g3.n, g3.m_decl, g3.src, g3.dst)
    op, oc, ip_, ic = E.download_csr()
    assert np.array_equal(op, g3.out_ptr) and np.array_equal(oc, g3.out_col) and np.array_equal(ip_, g3.in_ptr) and np.array_equal(ic, g3.in_col)
    # and the engine runs on it
    rmax, _ = E.configure("fora", EPS)
    res, rsd, rsum, _ = E.push_only(3, rmax)
    assert abs(res.sum() + rsd.sum() - 1) < 1e-12
    with pytest.raises(fb.ForaError):
        E.build_graph_from_edges(10, 3, np.array([1, 2, 10], np.int32), np.array([2, 3, 1], np.int32))  # id >= n
    E.build_graph_from_edges(5, 0, np.empty(0, np.int32), np.empty(0, np.int32))  # empty edge list
    op, oc, ip_, ic = E.download_csr()
    assert op.tolist() == [0] * 6 and len(oc) == 0
    E.close()


# ------------------------------------------------------------------------------------ push
def test_push_matches_schedule_matched_oracle(g, eng):
    rmax, omega = eng.configure("fora", EPS)
    O = Oracle(g)
    O.init_state(-1.0, 0)
    for s in interesting_sources(g):
        res, rsd, rsum, st = eng.push_only(s, rmax)
        O.reset_counters()
        r2 = O.push_sync(s, rmax, 1, 0)
        a, b = O.fwd()
        c = O.counters()
        assert relerr(res, a) < PUSH_RTOL and relerr(rsd, b) < PUSH_RTOL
        assert abs(rsum - r2) < 1e-12
        assert (st["edges_pushed"], st["vertices_pushed"], st["push_levels"]) == (c["edges_pushed"], c["vertices_pushed"], c["push_levels"])
        assert ((res > 0) == (a > 0)).all() and ((rsd > 0) == (b > 0)).all()  # same touched sets


def test_push_invariants_shared_with_reference_fifo(g, eng):
    # what the reference's FIFO result (algo.h:954-1018) and ours must both satisfy
    rmax, _ = eng.configure("fora", EPS)
    O = Oracle(g)
    O.init_state(-1.0, 0)
    deg = np.maximum(g.deg, 1)
    for s in interesting_sources(g)[:3]:
        res, rsd, rsum, _ = eng.push_only(s, rmax)
        assert abs(res.sum() + rsd.sum() - 1.0) < 1e-12 and abs(rsd.sum() - rsum) < 1e-12
        nd = g.deg > 0
        assert (rsd[nd] / deg[nd] < rmax).all() and (rsd[~nd] == 0).all()
        rs_fifo = O.push_fifo(s, rmax)
        assert abs(rsum - rs_fifo) < 0.5 * max(rsum, rs_fifo)  # same order of magnitude of leftover mass
        # reserve + residue-weighted PPR == exact PPR: finish the push on the oracle and compare
        O.set_fwd(res, rsd)
        O.push_sync(s, 1e-14, 0, 1)
        assert np.abs(O.fwd()[0] - O.power_iteration(s, 200)).max() < 1e-9


def test_push_edge_cases(g, eng):
    eng.configure("fora", EPS)
    dang = int(np.flatnonzero(g.deg == 0)[0])
    res, rsd, rsum, st = eng.push_only(dang, 1e-6)       # algo.h:961-965
    assert rsum == 0.0 and res[dang] == 1.0 and res.sum() == 1.0 and rsd.sum() == 0.0 and st["edges_pushed"] == 0
    s = int(np.argmax(g.deg))
    res, rsd, rsum, st = eng.push_only(s, 10.0)          # huge rmax: the source is still pushed once (algo.h:973)
    # only the source and its dangling out-neighbours (x/0 = +inf >= rmax, algo.h:1012) are ever pushed
    n_dang_nb = len(set(int(u) for u in g.out_col[g.out_ptr[s]:g.out_ptr[s + 1]] if g.deg[u] == 0))
    assert st["vertices_pushed"] == 1 + n_dang_nb and st["edges_pushed"] == g.deg[s] and res[s] >= 0.2 - 1e-15
    O = Oracle(g)
    O.init_state(-1.0, 0)
    O.push_sync(s, 10.0, 1, 0)
    assert relerr(res, O.fwd()[0]) < PUSH_RTOL and relerr(rsd, O.fwd()[1]) < PUSH_RTOL
    with pytest.raises(fb.ForaError):
        eng.push_only(g.n, 1e-6)
    with pytest.raises(fb.ForaError):
        eng.push_only(-1, 1e-6)


def test_resumable_rounds_match_oracle(g, eng):
    rmax, _ = eng.configure("fora", EPS)
    O = Oracle(g)
    O.init_state(-1.0, 0)
    for s in interesting_sources(g)[:3]:
        eng.push_begin(s)
        O.push_topk_begin(s)
        for k in range(5):
            rm = rmax * 8 / 2 ** k
            res, rsd, rsum, _ = eng.push_round(rm)
            r2 = O.push_sync(s, rm, 0, 1)
            a, b = O.fwd()
            assert relerr(res, a) < PUSH_RTOL and relerr(rsd, b) < PUSH_RTOL and abs(rsum - r2) < 1e-12


def test_push_on_golden_graph(gold):
    # the reference's own FIFO output (golden) and the GPU output agree on the invariants, and on the
    # exact values for the sources where both schedules coincide (dangling source)
    n = int(gold["n"])
    g2 = Graph(n, gold["src"], gold["dst"], int(gold["m_decl"]))
    E = fb.Engine(0)
    E.upload_graph(n, g2.m_decl, g2.out_ptr, g2.out_col)
    E.configure("fora", float(gold["eps"]))
    rmax = float(gold["rmax"])
    O = Oracle(g2)
    for i, s in enumerate(gold["sources"]):
        res, rsd, rsum, _ = E.push_only(int(s), rmax)
        assert abs(res.sum() + rsd.sum() - 1) < 1e-12
        if g2.deg[s] == 0:
            assert np.array_equal(res, gold["fifo_reserve_%d" % i]) and rsum == float(gold["fifo_rsum_%d" % i])
        O.set_fwd(res, rsd)
        O.push_sync(int(s), 1e-15, 0, 1)
        assert np.abs(O.fwd()[0] - gold["power_%d" % i]).max() < 1e-8
    E.close()


# ------------------------------------------------------------------------------------ walks
def test_walks_distribution_and_length(g, eng):
    eng.configure("fora", EPS)
    O = Oracle(g, seed=99)
    N = 400000
    for nzh in (0, 1):
        for s in interesting_sources(g)[:3]:
            d, hops = eng.random_walks(s, N, nzh)
            a = np.bincount(d, minlength=g.n).astype(np.float64)
            O.reset_counters()
            b = np.bincount(O.walks(s, N, nzh), minlength=g.n).astype(np.float64)
            m = (a + b) >= 20
            chi2 = (((a - b)[m] ** 2) / (a + b)[m]).sum()
            dof = m.sum() - 1
            assert chi2 < dof + 6 * np.sqrt(2 * dof), (chi2, dof)
            h_or = O.counters()["walk_hops"] / N
            assert abs(hops / N - h_or) < 0.03  # E[hops] ~ 4 (5 when the first hop is forced)
    dang = int(np.flatnonzero(g.deg == 0)[0])
    d, hops = eng.random_walks(dang, 1000, 1)
    assert (d == dang).all() and hops == 0  # algo.h:127-129
    d, _ = eng.random_walks(0, 0, 0)
    assert len(d) == 0


def test_walk_plan_counts_bit_exact_and_guarantee(gold):
    # compute_ppr_with_fwdidx{,_opt} (query.h:255-413) on the push state the REFERENCE produced (golden)
    n = int(gold["n"])
    g2 = Graph(n, gold["src"], gold["dst"], int(gold["m_decl"]))
    E = fb.Engine(0, seed=5)
    E.upload_graph(n, g2.m_decl, g2.out_ptr, g2.out_col)
    O = Oracle(g2)
    for opt in (0, 1):
        rmax, omega = E.configure("fora", float(gold["eps"]), opt=opt)
        O.set_params(float(gold["eps"]), rmax, omega, opt=opt)
        O.init_state(-1.0, 0)
        for i in (0, 1, 2, 3):
            res, rsd, rsum = gold["fifo_reserve_%d" % i], gold["fifo_residue_%d" % i], float(gold["fifo_rsum_%d" % i])
            ppr, st = E.compute_ppr(res, rsd, rsum)
            O.set_fwd(res, rsd)
            keys, cnt, inc = O.walk_plan(rsum, opt)
            assert st["n_walks"] == int(cnt.sum()) and st["n_sources"] == int((cnt > 0).sum())
            assert abs(ppr.sum() - 1.0) < 1e-9
            exact = gold["power_%d" % i]
            big = exact >= 1.0 / n
            assert (np.abs(ppr[big] - exact[big]) / exact[big]).max() < float(gold["eps"])
    # rsum == 0 (dangling source): ppr = reserve, no walks (query.h:267-268)
    ppr, st = E.compute_ppr(gold["fifo_reserve_4"], gold["fifo_residue_4"], 0.0)
    assert np.array_equal(ppr, gold["fifo_reserve_4"]) and st["n_walks"] == 0
    E.close()


# ------------------------------------------------------------------------------------ queries
@pytest.mark.parametrize("opt,balanced", [(0, 0), (1, 0), (1, 1), (0, 1)])
def test_fora_query_guarantee(g, eng, opt, balanced):
    eng.configure("fora", EPS, opt=opt, balanced=balanced)
    srcs = np.array(interesting_sources(g) + [5, 123, 77, 4000], np.int32)
    ppr, stats, tm = eng.query_batch("fora", srcs)
    O = Oracle(g)
    bad = 0
    total = 0
    for i, s in enumerate(srcs):
        exact = O.power_iteration(int(s), 150)
        assert abs(ppr[i].sum() - 1.0) < 1e-9
        big = exact >= 1.0 / g.n
        rel = np.abs(ppr[i][big] - exact[big]) / exact[big]
        bad += int((rel > EPS).sum())
        total += int(big.sum())
        if g.deg[s] == 0:
            assert ppr[i][s] == 1.0 and stats[i]["n_walks"] == 0
    assert bad <= max(1, total // g.n)  # p_f = 1/n per (s,t) pair
    assert tm["kernel_launches"] > 0


def test_query_is_reproducible_and_slot_independent(g):
    srcs = np.array([0, 11, 5, 123, 77, 9, 4000], np.int32)
    out = []
    for slots in (1, 3, 8):
        E = fb.Engine(0, seed=21, slots=slots)
        E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col)
        E.configure("fora", EPS, opt=1)
        ppr, stats, _ = E.query_batch("fora", srcs)
        out.append((ppr, [s["n_walks"] for s in stats], [s["walk_hops"] for s in stats]))
        E.close()
    for ppr, nw, hops in out[1:]:
        assert nw == out[0][1] and hops == out[0][2]         # Philox keyed by (seed, query, source, walk)
        assert np.allclose(ppr, out[0][0], rtol=1e-9, atol=1e-15)
    E = fb.Engine(0, seed=22, slots=3)
    E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col)
    E.configure("fora", EPS, opt=1)
    ppr2, stats2, _ = E.query_batch("fora", srcs)
    assert [s["walk_hops"] for s in stats2] != out[0][2]  # another seed, other walks
    E.close()


def test_balanced_loop_matches_oracle_cost_model(g, eng):
    # query.h:848-884 with the wall clock replaced by the device cost model: rounds, final rmax
    # and the push state agree with the oracle running the same model on the same schedule
    cost = dict(cost_walk=3e-9, cost_edge=2e-9, cost_vertex=1e-9, cost_level=1e-7)
    rmax, omega = eng.configure("fora", EPS, opt=1, balanced=1, **cost)
    srcs = np.array([0, 11, 5, int(np.argmax(g.deg))], np.int32)
    _, stats, _ = eng.query_batch("fora", srcs, want_ppr=False)
    O = Oracle(g)
    O.set_params(EPS, rmax, omega, opt=1, balanced=1)
    O.init_state(-1.0, 0)
    # walks are irrelevant here: give the oracle a tiny omega for the walk phase only
    for i, s in enumerate(srcs):
        O.reset_counters()
        O.set_params(EPS, rmax, omega, opt=1, balanced=1)
        O.push_topk_begin(int(s))
        rm, used, rsum, rounds = rmax * 8, 0.0, 1.0, 0
        while omega * rsum * 0.8 * cost["cost_walk"] > used:
            c0 = O.counters()
            rsum = O.push_sync(int(s), rm, 0, 1)
            c1 = O.counters()
            used += cost["cost_edge"] * (c1["edges_pushed"] - c0["edges_pushed"]) + cost["cost_vertex"] * (
                c1["vertices_pushed"] - c0["vertices_pushed"]) + cost["cost_level"] * (c1["push_levels"] - c0["push_levels"])
            rm /= 2
            rounds += 1
        assert stats[i]["push_rounds"] == rounds and stats[i]["final_rmax"] == rm * 2
        assert abs(stats[i]["rsum"] - rsum) < 1e-12
        assert stats[i]["edges_pushed"] == O.counters()["edges_pushed"]


def test_montecarlo_and_fwdpush(g, eng):
    O = Oracle(g)
    s = 11
    exact = O.power_iteration(s, 150)
    big = exact >= 1.0 / g.n
    rmax, omega = eng.configure("montecarlo", EPS)
    ppr, stats, _ = eng.query_batch("montecarlo", np.array([s], np.int32))
    assert stats[0]["n_walks"] == int(np.ceil(omega)) and abs(ppr[0].sum() - np.ceil(omega) / omega) < 1e-9
    assert (np.abs(ppr[0][big] - exact[big]) / exact[big]).max() < EPS
    # fwdpush: ppr == reserve of a push at rmax = eps/m-ish (query.h:1503-1508); schedule-matched oracle
    rmax, _ = eng.configure("fwdpush", EPS)
    ppr, stats, _ = eng.query_batch("fwdpush", np.array([s, 0], np.int32))
    O.init_state(-1.0, 0)
    for i, src in enumerate((s, 0)):
        O.push_sync(src, rmax, 1, 0)
        assert relerr(ppr[i], O.fwd()[0]) < PUSH_RTOL
    # pure forward push only under-estimates (reserve <= pi; algo.h:486 "no accuracy guarantee")
    assert (ppr[0] <= exact + 1e-12).all() and np.average(np.abs(ppr[0][big] - exact[big]) / exact[big]) < 0.5


def test_reverse_push_matches_oracle(g, eng, gold):
    # reverse_local_update_linear (algo.h:703-751) on the frontier-synchronous schedule
    eng.configure("bippr", EPS)
    O = Oracle(g)
    for rm in (0.3, 1e-2, 1e-4, 2.0):
        for t in interesting_sources(g):
            res, rsd = eng.reverse_push(t, rm)
            O.reverse_push(t, rm, 1.0, 1)
            a, b = O.bwd()
            assert relerr(res, a) < PUSH_RTOL and relerr(rsd, b) < PUSH_RTOL, (rm, t)
    # on the golden graph a shallow push (rmax = 0.3) visits every vertex at most once, so the reference's
    # FIFO result (golden) and the synchronous schedule coincide up to summation order
    n = int(gold["n"])
    g2 = Graph(n, gold["src"], gold["dst"], int(gold["m_decl"]))
    E = fb.Engine(0)
    E.upload_graph(n, g2.m_decl, g2.out_ptr, g2.out_col, g2.in_ptr, g2.in_col)
    E.configure("bippr", float(gold["eps"]))
    res, rsd = E.reverse_push(int(gold["sources"][1]), float(gold["bwd_rmax_0"]))
    assert np.allclose(res, gold["bwd_reserve_0"], rtol=1e-12, atol=0) and np.allclose(rsd, gold["bwd_residue_0"], rtol=1e-12, atol=0)
    E2 = fb.Engine(0)
    E2.upload_graph(n, g2.m_decl, g2.out_ptr, g2.out_col)  # no in-CSR
    E2.configure("bippr", float(gold["eps"]))
    with pytest.raises(fb.ForaError, match="in-CSR"):
        E2.reverse_push(0, 0.3)
    E.close(); E2.close()


def test_bippr_query(g, eng):
    # bippr_query (query.h:71-124): omega walks + one backward push per target node
    rmax, omega = eng.configure("bippr", EPS)
    assert rmax < 1.0
    O = Oracle(g)
    srcs = np.array([11, 0], np.int32)
    ppr, stats, _ = eng.query_batch("bippr", srcs)
    for i, s in enumerate(srcs):
        exact = O.power_iteration(int(s), 150)
        big = exact >= 1.0 / g.n
        assert stats[i]["n_walks"] == int(np.ceil(omega))
        assert (np.abs(ppr[i][big] - exact[big]) / exact[big]).max() < 1.0
        assert np.median(np.abs(ppr[i][big] - exact[big]) / exact[big]) < 0.15
    # rmax >= 1 falls back to pure Monte-Carlo (query.h:114-120)
    eng.set_params(EPS, 1.5, omega)
    ppr2, st2, _ = eng.query_batch("bippr", srcs[:1])
    assert abs(ppr2[0].sum() - np.ceil(omega) / omega) < 1e-9


# ------------------------------------------------------------------------------------ index
def test_index_layout_bit_exact_and_with_idx_queries(gold):
    n = int(gold["n"])
    g2 = Graph(n, gold["src"], gold["dst"], int(gold["m_decl"]))
    for opt in (0, 1):
        E = fb.Engine(0, seed=3)
        E.upload_graph(n, g2.m_decl, g2.out_ptr, g2.out_col)
        rmax, omega = E.configure("fora", float(gold["eps"]), opt=opt)
        off, cnt, total = E.index_info()
        assert np.array_equal(off, gold["idx_off_opt%d" % opt]) and np.array_equal(cnt, gold["idx_cnt_opt%d" % opt])
        dest = E.index_build(off, cnt)
        assert len(dest) == total and dest.min() >= 0 and dest.max() < n
        # sharded build (multi-GPU partitioning by source range) produces the same destinations
        mid = n // 3
        d2 = np.concatenate([E.index_build(off, cnt, 0, mid), E.index_build(off, cnt, mid, n)])
        assert np.array_equal(d2, dest)
        # per-source destination distribution vs oracle walks (one hub source)
        O = Oracle(g2, seed=17)
        O.set_params(float(gold["eps"]), rmax, omega, opt=opt)
        # with_idx query
        E.index_upload(off, cnt, dest)
        E.configure("fora", float(gold["eps"]), opt=opt, with_idx=1)
        srcs = gold["sources"][:4].astype(np.int32)
        ppr, stats, _ = E.query_batch("fora", srcs)
        for i, s in enumerate(srcs):
            exact = gold["power_%d" % i]
            big = exact >= 1.0 / n
            assert (np.abs(ppr[i][big] - exact[big]) / exact[big]).max() < float(gold["eps"])
            assert stats[i]["n_idx_hits"] > 0.9 * stats[i]["n_walks"]
            # hit accounting == the reference's rule min(n_v, count_v) summed over sources (query.h:290-307)
            res, rsd, rsum, _ = E.push_only(int(s), rmax)
            O.set_fwd(res, rsd)
            keys, c, inc = O.walk_plan(rsum, opt)
            assert stats[i]["n_idx_hits"] == int(np.minimum(c, cnt[keys]).sum())
        E.close()


# ------------------------------------------------------------------------------------ top-k / ground truth
def test_topk_select(g, eng):
    rng = np.random.default_rng(0)
    v = np.zeros(g.n)
    idx = rng.choice(g.n, 3000, replace=False)
    v[idx] = rng.random(3000) ** 4
    v[idx[:500]] = 0.125  # a large tie group straddling the k-th position for some k
    for k in (1, 7, 500, 1000, 2999, 3000, 3500):
        nodes, vals = eng.topk_of(v, k)
        order = np.lexsort((np.arange(g.n), -v))[:k]
        exp_nodes = np.where(v[order] > 0, order, 0)
        exp_vals = v[order]
        assert np.array_equal(vals, exp_vals) and np.array_equal(nodes, exp_nodes), k


@pytest.mark.parametrize("with_idx", [0, 1])
def test_topk_batch_fora_opt(with_idx):
    # fora_query_topk_new (query.h:972-1045): precision against power iteration, and against the oracle's
    # restatement of the same driver (north_star: within 0.01 of the reference on average; small graph here)
    g2 = Graph.synth(6000, 72000, seed=13)
    k = 50
    E = fb.Engine(0, seed=4, slots=3)
    E.upload_graph(g2.n, g2.m_decl, g2.out_ptr, g2.out_col)
    rmax, omega = E.configure("fora", EPS, opt=1, k=k, with_idx=with_idx)
    if with_idx:
        off, cnt, total = E.index_info()
        E.index_upload(off, cnt, E.index_build(off, cnt))
    srcs = np.array([21, 5, 300, int(np.argmax(g2.deg)), int(np.flatnonzero(g2.deg == 0)[0]), 77, 1234], np.int32)
    nodes, vals, iters, stats, tm = E.topk_batch("fora", srcs, k)
    O = Oracle(g2, seed=8)
    O.set_params(EPS, 0.0, 0.0, opt=1, k=k, with_idx=with_idx)
    O.init_state(-9.0, 1)
    if with_idx:
        O.index_set(off, cnt, E.index_build(off, cnt))
    prec_gpu, prec_orc = [], []
    for i, s in enumerate(srcs):
        exact = O.power_iteration(int(s), 150)
        top = np.argsort(-exact, kind="stable")[:k]
        npos = int((exact > 0).sum())
        assert (np.diff(vals[i]) <= 0).all()
        if g2.deg[s] == 0:
            assert nodes[i][0] == s and vals[i][0] == 1.0 and (vals[i][1:] == 0).all() and iters[i] >= 1
            continue
        assert iters[i] >= 1 and stats[i]["n_walks"] > 0
        kk = min(k, npos)
        prec_gpu.append(len(set(nodes[i][:kk].tolist()) & set(top[:kk].tolist())) / kk)
        O.fora_topk_new(int(s), 0)
        on, ov = O.topk_ppr(k)
        prec_orc.append(len(set(on[:kk].tolist()) & set(top[:kk].tolist())) / kk)
        if with_idx:
            assert stats[i]["n_idx_hits"] > 0
    assert np.mean(prec_gpu) > 0.9 and abs(np.mean(prec_gpu) - np.mean(prec_orc)) < 0.05, (prec_gpu, prec_orc)
    assert tm["topk_ms"] > 0
    with pytest.raises(fb.ForaError):
        E.topk_batch("fora", srcs, 1)            # 1 < k < n-1 (query.h:1317-1318)
    E.close()


def test_topk_batch_fora_with_bound():
    # fora_query_topk_with_bound (query.h:909-969) + set_ppr_bounds / if_stop (algo.h:1096-1261): non --opt top-k
    g2 = Graph.synth(6000, 72000, seed=13)
    k = 50
    E = fb.Engine(0, seed=4, slots=4)
    E.upload_graph(g2.n, g2.m_decl, g2.out_ptr, g2.out_col)
    E.configure("fora", EPS, opt=0, k=k)
    srcs = np.array([21, 5, 300, int(np.flatnonzero(g2.deg == 0)[0]), 77], np.int32)
    nodes, vals, iters, stats, tm = E.topk_batch("fora", srcs, k)
    O = Oracle(g2, seed=8)
    O.set_params(EPS, 0.0, 0.0, opt=0, k=k)
    O.init_state(-9.0, 1)
    pg, po, itg, ito = [], [], [], []
    for i, s in enumerate(srcs):
        if g2.deg[s] == 0:
            assert nodes[i][0] == s and vals[i][0] == 1.0
            continue
        exact = O.power_iteration(int(s), 150)
        top = set(np.argsort(-exact, kind="stable")[:k].tolist())
        pg.append(len(set(nodes[i].tolist()) & top) / k)
        O.reset_counters()
        O.fora_topk_with_bound(int(s), 0)
        on, ov = O.topk_ppr(k)
        po.append(len(set(on.tolist()) & top) / k)
        itg.append(int(iters[i])); ito.append(O.counters()["topk_iters"])
        assert (np.diff(vals[i]) <= 0).all()
    assert np.mean(pg) > 0.9 and abs(np.mean(pg) - np.mean(po)) < 0.05, (pg, po)
    assert abs(np.mean(itg) - np.mean(ito)) <= 2.0, (itg, ito)  # same stopping rule => similar number of rounds
    E.close()


def test_topk_batch_baselines(g, eng):
    k = 30
    O = Oracle(g)
    s = 11
    exact = O.power_iteration(s, 150)
    top = set(np.argsort(-exact, kind="stable")[:k].tolist())
    for algo in ("montecarlo", "fwdpush"):
        eng.configure(algo, EPS, k=k)
        nodes, vals, iters, stats, tm = eng.topk_batch(algo, np.array([s], np.int32), k)
        assert len(set(nodes[0].tolist()) & top) / k > 0.85 and (np.diff(vals[0]) <= 0).all()


def test_power_iteration_matches_golden(gold):
    n = int(gold["n"])
    g2 = Graph(n, gold["src"], gold["dst"], int(gold["m_decl"]))
    E = fb.Engine(0)
    E.upload_graph(n, g2.m_decl, g2.out_ptr, g2.out_col)
    E.configure("fora", 0.5)
    for i, s in enumerate(gold["sources"]):
        assert np.abs(E.power_iteration(int(s), 100) - gold["power_%d" % i]).max() < 1e-13
    E.close()


# ------------------------------------------------------------------------------------ full size
def test_livejournal_shape_properties():
    # BASELINE.json config 2 shape (4.8M nodes / 69M edges): size-independent properties only
    n, m = 4847571, 68993773
    src, dst = fb.synth_edges(n, m, 42)
    op, oc, _, _ = fb.csr_from_edges(n, src, dst, with_in=False)
    del src, dst
    deg = np.diff(op)
    E = fb.Engine(0, seed=1, slots=2)
    E.upload_graph(n, m, op, oc)
    rmax, omega = E.configure("fora", EPS, opt=1, balanced=1)
    s = int(np.flatnonzero(deg > 5)[12345])
    res, rsd, rsum, st = E.push_only(s, rmax)
    assert abs(res.sum() + rsd.sum() - 1.0) < 1e-11 and abs(rsd.sum() - rsum) < 1e-12
    nd = deg > 0
    assert (rsd[nd] < rmax * deg[nd]).all() and (rsd[~nd] == 0).all()
    assert st["edges_pushed"] <= 1.0 / (0.2 * rmax)  # work bound sum d_out <= 1/(alpha*rmax)
    srcs = np.array([s, int(np.flatnonzero(deg == 0)[7]), 17], np.int32)
    ppr, stats, tm = E.query_batch("fora", srcs)
    for i in range(3):
        assert abs(ppr[i].sum() - 1.0) < 1e-9 and ppr[i].min() >= 0.0
    assert ppr[1][srcs[1]] == 1.0
    # linearity of the estimator in expectation: the hop count per online walk is ~5 (no-zero-hop)
    assert 4.0 < stats[0]["walk_hops"] / max(stats[0]["n_walks"], 1) < 5.2
    # top-k of the estimate contains the source
    nodes, vals = E.topk_of(ppr[0], 500)
    assert nodes[0] == s and (np.diff(vals) <= 0).all() and vals[-1] > 0
    E.close()
