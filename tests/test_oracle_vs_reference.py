"""Pins the CPU oracle (oracle/fora_oracle.c) against the UNMODIFIED reference compiled by
oracle/Makefile (oracle/_ref/libfora_ref.so, Boost replaced by oracle/boost_shim).

The reference ships no tests or golden vectors, so this file is where parity
of the restatement is established: bit-exact for integer / deterministic double work, statistical
for anything driven by the reference's time(0)-seeded RNGs.  Skipped when oracle/_ref has not been
built (the golden fixtures under tests/golden/ cover the same ground without it).
"""
import os
import tempfile

import numpy as np
import pytest

from helpers import Graph, Oracle, Reference, have_reference, write_dataset

pytestmark = pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")

EPS = 0.5


@pytest.fixture(scope="module")
def g():
    return Graph.synth(3000, 30000, seed=7, self_loops=20)


def sources(g):
    return [0, 5, 17, int(np.argmax(g.deg)), int(np.flatnonzero(g.deg == 0)[0])]


def test_text_loader_and_csr(g):
    # graph.h:48-64,152-160: self loops dropped, duplicates and file order kept, m taken from attribute.txt
    d = tempfile.mkdtemp()
    write_dataset(d, g.n, g.m_decl, g.src, g.dst)
    R = Reference(folder=d + "/", epsilon=EPS)
    op, oc, ip_, ic = R.graph_dump()
    assert R.lib.ref_graph_m() == g.m_decl
    assert np.array_equal(op, g.out_ptr) and np.array_equal(oc, g.out_col)
    assert np.array_equal(ip_, g.in_ptr) and np.array_equal(ic, g.in_col)
    assert (g.src == g.dst).sum() == 20 and len(oc) == len(g.src) - 20
    # the oracle's own loader restatement
    import ctypes as C
    O = Oracle(g)
    n, m = C.c_int(0), C.c_longlong(0)
    assert O.lib.orc_read_attribute(os.path.join(d, "attribute.txt").encode(), C.byref(n), C.byref(m)) == 0
    assert (n.value, m.value) == (g.n, g.m_decl)
    O.lib.orc_read_edges.restype = C.c_longlong
    cnt = O.lib.orc_read_edges(os.path.join(d, "graph.txt").encode(), g.n, None, None)
    assert cnt == len(oc)


@pytest.mark.parametrize("opt", [0, 1])
def test_settings_bit_exact(g, opt):
    # algo.h:442-496
    R = Reference(g, epsilon=EPS, opt=opt, rmax_scale=1.0)
    O = Oracle(g)
    for w in ["fora", "fora_topk", "bippr", "fwdpush"]:
        assert R.setting(w) [0] == O.setting(w, EPS, opt=opt)[0]
    for w in ["fora", "fora_topk", "montecarlo", "bippr"]:
        assert R.setting(w)[1] == O.setting(w, EPS, opt=opt)[1]


def test_forward_push_fifo_bit_exact(g):
    # algo.h:954-1018: same sequential order => identical doubles and identical insertion order
    R = Reference(g, epsilon=EPS)
    O = Oracle(g)
    rmax, omega = R.setting("fora")
    R.init_query_state()
    O.init_state(-1.0, 0)
    for s in sources(g):
        r1 = R.push(s, rmax)
        a, b, ao, bo = R.fwd()
        r2 = O.push_fifo(s, rmax)
        c, d = O.fwd()
        assert r1 == r2
        assert np.array_equal(a, c) and np.array_equal(b, d)
        assert np.array_equal(ao, O.reserve_occur()) and np.array_equal(bo, O.residue_occur())


def test_resumable_push_bit_exact(g):
    # algo.h:1020-1093 over the --balanced schedule 8*rmax, 4*rmax, ... (query.h:863-877)
    R = Reference(g, epsilon=EPS)
    O = Oracle(g)
    rmax, _ = R.setting("fora")
    R.init_query_state()
    O.init_state(-1.0, 0)
    for s in sources(g)[:4]:
        R.push_topk_begin(s)
        O.push_topk_begin(s)
        for k in range(6):
            rm = rmax * 8 / 2 ** k
            r1 = R.push_topk_round(s, rm, rmax)
            r2 = O.push_topk_round(s, rm, rmax)
            a, b, _, _ = R.fwd()
            c, d = O.fwd()
            assert r1 == r2 and np.array_equal(a, c) and np.array_equal(b, d)
            assert np.array_equal(R.push_topk_candidates(), O.push_topk_candidates())


def test_reverse_push_bit_exact(g):
    # algo.h:703-751 (including its early break)
    R = Reference(g, epsilon=EPS)
    O = Oracle(g)
    R.init_query_state()
    for rm in [0.3, 0.01, 1e-4]:
        R.lib.ref_set_rmax_omega(rm, 1000.0)
        for t in sources(g):
            R.reverse_push(t)
            a, b = R.bwd()
            O.reverse_push(t, rm, 1.0, 0)
            c, d = O.bwd()
            assert np.array_equal(a, c) and np.array_equal(b, d)


def test_sync_push_same_invariants_as_fifo(g):
    # The frontier-synchronous schedule (the CUDA kernel's) is a different but equally valid push:
    # mass is conserved, every residue ends below rmax*d_out, and reserve + sum_v residue(v)*pi(v,.)
    # reproduces the exact PPR vector, exactly as for the reference's FIFO result.
    O = Oracle(g)
    rmax, omega = O.setting("fora", EPS)
    O.init_state(-1.0, 0)
    deg = np.maximum(g.deg, 1)
    for s in sources(g)[:4]:
        exact = O.power_iteration(s, 200)
        outs = []
        for sync in (0, 1):
            rs = O.push_sync(s, rmax, 1, 0) if sync else O.push_fifo(s, rmax)
            res, rsd = O.fwd()
            assert abs(res.sum() + rsd.sum() - 1.0) < 1e-12
            assert abs(rsd.sum() - rs) < 1e-12
            nd = g.deg > 0
            assert (rsd[nd] / deg[nd] < rmax).all()
            assert (rsd[~nd] == 0).all()  # dangling vertices never keep residue
            outs.append((res, rsd))
        # pi_s = reserve + sum_v residue_v * pi_v: continuing the push from the synchronous state to a
        # vanishing rmax must converge to the exact vector
        O.set_fwd(outs[1][0], outs[1][1])
        O.push_sync(s, 1e-14, 0, 1)
        res2, rsd2 = O.fwd()
        assert np.abs(res2 - exact).max() < 1e-9


def test_walk_destination_distribution(g):
    # algo.h:124-166: oracle walks vs reference walks, chi-square on destination histograms
    R = Reference(g, epsilon=EPS)
    O = Oracle(g, seed=123)
    N = 200000
    for nzh in (0, 1):
        for s in sources(g)[:3]:
            a = np.bincount(R.walks(s, N, nzh), minlength=g.n).astype(np.float64)
            b = np.bincount(O.walks(s, N, nzh), minlength=g.n).astype(np.float64)
            m = (a + b) >= 20
            chi2 = (((a - b)[m] ** 2) / (a + b)[m]).sum()
            dof = m.sum() - 1
            assert chi2 < dof + 6 * np.sqrt(2 * dof), (chi2, dof)
    # dangling start returns itself (algo.h:127-129)
    dang = int(np.flatnonzero(g.deg == 0)[0])
    assert (R.walks(dang, 100, 0) == dang).all() and (O.walks(dang, 100, 1) == dang).all()


@pytest.mark.parametrize("opt", [0, 1])
def test_walk_plan_and_ppr(g, opt):
    # query.h:255-413: per-source walk counts / increments are pure arithmetic on the push state;
    # compared through the reference by replaying its loop with the state it produced
    R = Reference(g, epsilon=EPS, opt=opt)
    O = Oracle(g, seed=5)
    rmax, omega = R.setting("fora")
    O.set_params(EPS, rmax, omega, opt=opt)
    R.init_query_state()
    O.init_state(-1.0, 0)
    s = 17
    rsum = R.push(s, rmax)
    assert O.push_fifo(s, rmax) == rsum
    R.lib.ref_reset_counters()
    R.compute_ppr("opt" if opt else "fwdidx", rsum)
    total_ref, _ = R.counters()
    keys, cnt, inc = O.walk_plan(rsum, opt)
    assert int(cnt.sum()) == total_ref  # num_total_rw
    O.compute_ppr("opt" if opt else "fwdidx", rsum)
    p_ref, p_orc = R.ppr(), O.ppr()
    assert abs(p_ref.sum() - 1.0) < 1e-9 and abs(p_orc.sum() - 1.0) < 1e-9
    exact = O.power_iteration(s, 200)
    big = exact >= 1.0 / g.n
    for p in (p_ref, p_orc):
        assert (np.abs(p[big] - exact[big]) / exact[big]).max() < EPS


def test_index_info_bit_exact_and_queries(g):
    # build.h:302-366: offsets / counts are a pure function of degrees, rmax, omega (bit-exact);
    # destinations are random.  Then --with_idx queries (query.h:277-309).
    for opt in (0, 1):
        R = Reference(g, epsilon=EPS, opt=opt)
        O = Oracle(g, seed=9)
        rmax, omega = R.setting("fora")
        O.set_params(EPS, rmax, omega, opt=opt)
        d = tempfile.mkdtemp()
        off, cnt, dest = R.build_index(d)
        o2, c2, total = O.index_info()
        assert np.array_equal(off, o2) and np.array_equal(cnt, c2) and total == len(dest)
        names = (R.lib.ref_index_file_names(0).decode(), R.lib.ref_index_file_names(1).decode())
        assert names[0].endswith("randwalks.idx.onehopopt" if opt else "randwalks.idx")
        assert os.path.exists(names[0]) and os.path.exists(names[1])
        # reference round trip through its (shimmed) archives
        R2 = Reference(g, epsilon=EPS, opt=opt, with_idx=1)
        R2.setting("fora")
        R2.lib.ref_load_index((d + "/").encode())
        off3, cnt3 = np.zeros(g.n, np.uint64), np.zeros(g.n, np.uint64)
        dest3 = np.zeros(len(dest), np.int32)
        from helpers import _p, c_up, c_ip
        R2.lib.ref_index_dump(_p(off3, c_up), _p(cnt3, c_up), _p(dest3, c_ip))
        assert np.array_equal(off3, off) and np.array_equal(cnt3, cnt) and np.array_equal(dest3, dest)
        # query with index: oracle vs exact
        O.set_params(EPS, rmax, omega, opt=opt, with_idx=1)
        O.init_state(-1.0, 0)
        O.index_set(off, cnt, dest)
        s = 5
        rsum = O.push_fifo(s, rmax)
        O.reset_counters()
        O.compute_ppr("opt" if opt else "fwdidx", rsum)
        c = O.counters()
        assert c["hit_idx"] > 0.9 * c["total_rw"]
        R2.init_query_state()
        rs2 = R2.push(s, rmax)
        R2.lib.ref_reset_counters()
        R2.compute_ppr("opt" if opt else "fwdidx", rs2)
        tot, hit = R2.counters()
        assert (tot, hit) == (c["total_rw"], c["hit_idx"])
        exact = O.power_iteration(s, 200)
        big = exact >= 1.0 / g.n
        for p in (O.ppr(), R2.ppr()):
            assert (np.abs(p[big] - exact[big]) / exact[big]).max() < EPS


def test_power_iteration(g):
    # query.h:1192-1224 (100 sweeps); unordered_map iteration order only changes rounding
    R = Reference(g, epsilon=EPS)
    O = Oracle(g)
    for s in sources(g)[:3]:
        a, b = R.power_iteration(s), O.power_iteration(s, 100)
        assert np.abs(a - b).max() < 1e-13
        assert abs(a.sum() - (1 - 0.8 ** 100)) < 1e-9


def test_topk_and_precision(g):
    # algo.h:524-572,592-610
    R = Reference(g, epsilon=EPS, k=20)
    O = Oracle(g)
    exact = O.power_iteration(3, 100)
    order = np.argsort(-exact, kind="stable")[:20]
    est_nodes = np.r_[order[:12], np.arange(2000, 2008)].astype(np.int32)
    est_vals = np.r_[exact[order[:12]], np.full(6, 1e-9), [0.0, 0.0]]
    p1, r1 = R.precision(3, est_nodes, est_vals, order.astype(np.int32), exact[order])
    p2, r2 = O.precision(20, est_nodes, est_vals, order.astype(np.int32), exact[order])
    assert (p1, r1) == (p2, r2) and abs(p1 - 12 / 20) < 1e-12


@pytest.mark.parametrize("algo", ["fora", "montecarlo", "fwdpush", "bippr"])
def test_query_drivers_agree_with_exact(algo):
    # whole drivers (query.h:16-193, 841-907, 1503-1508) on a smaller graph: oracle and reference both
    # satisfy the (eps, delta) guarantee against the power-iteration ground truth
    g = Graph.synth(600, 5000, seed=11)
    eps = 0.5
    R = Reference(g, epsilon=eps)
    O = Oracle(g, seed=77)
    rmax, omega = R.setting(algo)
    O.set_params(eps, rmax, omega)
    R.init_query_state()
    O.init_state(-1.0, 0)
    if algo in ("montecarlo", "bippr"):
        O.init_state(0.0, 0)
    s = 9
    exact = O.power_iteration(s, 200)
    big = exact >= 1.0 / g.n
    R.query(algo, s)
    if algo == "fora":
        O.fora_query(s)
    elif algo == "montecarlo":
        O.montecarlo_query(s)
    elif algo == "fwdpush":
        O.fwdpush_query(s)
    else:
        O.bippr_query(s)
    tol = {"fora": eps, "montecarlo": eps, "fwdpush": 1.0, "bippr": 1.0}[algo]
    for p in (R.ppr(), O.ppr()):
        assert (np.abs(p[big] - exact[big]) / exact[big]).max() < tol
    if algo == "fwdpush":
        assert np.array_equal(R.ppr(), O.ppr())  # deterministic: bit-exact


@pytest.mark.parametrize("opt", [0, 1])
def test_topk_drivers(opt):
    # fora_query_topk_new (--opt, query.h:972-1045) / fora_query_topk_with_bound (query.h:909-969): average top-k precision
    # (the reference's compute_precision, algo.h:524-572) of the oracle within 0.01 of the unmodified reference's -- north_star's
    # tolerance -- over 40 queries at k=200 (8000 ranked entries per arm: the sampling error of the difference is ~0.003)
    g = Graph.synth(20000, 240000, seed=13)
    k, nq = 200, 40
    R = Reference(g, epsilon=0.5, opt=opt, k=k)
    O = Oracle(g, seed=3)
    O.set_params(0.5, 0.0, 0.0, opt=opt, k=k)
    R.init_topk_state("fora")
    O.init_state(-9.0, 1)
    srcs = np.random.default_rng(5).choice(np.flatnonzero(g.deg > 0), nq, replace=False)
    pr, po = [], []
    for s in srcs:
        s = int(s)
        exact = O.power_iteration(s, 100)
        top = np.argsort(-exact, kind="stable")[:k].astype(np.int32)
        nodes_r, vals_r = R.topk("fora", s, k)
        if opt:
            O.fora_topk_new(s)
        else:
            O.fora_topk_with_bound(s)
        nodes_o, vals_o = O.topk_ppr(k)
        pr.append(R.precision(s, nodes_r, vals_r, top, exact[top])[0])
        po.append(R.precision(s, nodes_o, vals_o, top, exact[top])[0])
        assert (np.diff(vals_o) <= 0).all() and (np.diff(vals_r) <= 0).all()
    assert np.mean(pr) >= 0.9 and np.mean(po) >= 0.9 and abs(np.mean(pr) - np.mean(po)) <= 0.01, (np.mean(pr), np.mean(po))
