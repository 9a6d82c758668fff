"""GPU accuracy pins at the sizes and tolerances north_star names (round 2, VERDICT "Next #2"):

(i)  the headline bench configuration itself -- LiveJournal-shape graph, FORA eps=0.5 ``--balanced --opt`` with the
     engine's built-in cost calibration -- against exact PPR (``fora_power_iteration``, itself pinned at 1e-13 against the
     reference's power iteration in test_gpu_parity.py): FORA's (eps, delta=1/n, p_f=1/n) guarantee on every pi >= 1/n
     (/root/reference/query.h:1192-1224 is the ground truth, query.h:821-907 the query).
(ii) top-k precision of both FORA top-k drivers (query.h:909-1045) averaged over 100 queries of a 100K-node graph at
     k=500, against the UNMODIFIED reference (oracle/_ref) run on the same queries -- |delta| <= 0.01 (north_star) with the
     reference's own precision definition (algo.h:524-572).  Where oracle/_ref has not been built the C oracle stands in.
"""
import numpy as np
import pytest

import fora_b200 as fb
from helpers import Graph, Oracle, Reference, have_reference

pytestmark = pytest.mark.gpu
EPS = 0.5
TOPK_TOL = 0.01  # north_star: "top-k precision must match the reference's within 0.01"


@pytest.mark.parametrize("shared", [0, 1])  # 1: the opt-in per-wave walk pool (fora_ctx_set_shared_walks) at the same scale
def test_livejournal_shape_accuracy_vs_exact_ppr(shared):
    n, m = 4847571, 68993773
    src, dst = fb.synth_edges(n, m, 42)
    op, oc, _, _ = fb.csr_from_edges(n, src, dst, with_in=False)
    del src, dst
    deg = np.diff(op)
    E = fb.Engine(0, seed=2026, slots=12)
    E.upload_graph(n, m, op, oc)
    E.configure("fora", EPS, opt=1, balanced=1)  # the bench configuration: built-in --balanced calibration
    E.set_shared_walks(bool(shared))
    rng = np.random.default_rng(43)              # the bench's query list (seed 43), first ids + structural extremes
    srcs = np.r_[rng.integers(0, n, 1000)[:9], [int(np.argmax(deg)), int(np.flatnonzero(deg == 1)[3]), int(np.flatnonzero(deg == 0)[5])]].astype(np.int32)
    ppr, stats, _ = E.query_batch("fora", srcs)
    bad = total = 0
    worst = 0.0
    for i, s in enumerate(srcs):
        exact = E.power_iteration(int(s), 100)
        assert abs(ppr[i].sum() - 1.0) < 1e-9 and ppr[i].min() >= 0.0
        big = exact >= 1.0 / n
        rel = np.abs(ppr[i][big] - exact[big]) / exact[big]
        bad += int((rel > EPS).sum())
        total += int(big.sum())
        worst = max(worst, float(rel.max()) if rel.size else 0.0)
        if deg[s] == 0:
            assert ppr[i][s] == 1.0
        else:
            assert stats[i]["push_rounds"] >= 1 and stats[i]["n_walks"] > 0
    # p_f = 1/n per (source, target) pair: the expected number of violations over `total` pairs is total/n << 1
    assert total > 10000 and bad <= max(1, total // n), (bad, total, worst)
    E.close()


@pytest.mark.parametrize("opt", [1, 0])
def test_topk_precision_matches_reference_k500(opt):
    g = Graph.synth(100000, 1200000, seed=31)
    k, nq = 500, 100
    srcs = np.random.default_rng(7).choice(np.flatnonzero(g.deg > 0), nq, replace=False).astype(np.int32)
    E = fb.Engine(0, seed=99, slots=25)
    E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col)
    E.configure("fora", EPS, opt=opt, k=k)
    nodes, vals, iters, stats, _ = E.topk_batch("fora", srcs, k)
    exact_top = []
    for s in srcs:
        ex = E.power_iteration(int(s), 100)  # gen_exact_topk keeps the k largest (query.h:1226-1307)
        order = np.argsort(-ex, kind="stable")[:k].astype(np.int32)
        exact_top.append((order, ex[order]))
    E.close()
    O = Oracle(g, seed=5)
    if have_reference():
        R = Reference(g, epsilon=EPS, opt=opt, k=k)
        R.init_topk_state("fora")
        chk = R
    else:  # the GPU box always carries oracle/_ref; this branch keeps the test meaningful without it
        O.set_params(EPS, 0.0, 0.0, opt=opt, k=k)
        O.init_state(-9.0, 1)
        chk = None
    p_gpu, p_ref = [], []
    for i, s in enumerate(srcs):
        en, ev = exact_top[i]
        assert (np.diff(vals[i]) <= 0).all() and iters[i] >= 1
        if chk is not None:
            rn, rv = chk.topk("fora", int(s), k)
            p_ref.append(chk.precision(int(s), rn, rv, en, ev)[0])       # the reference's own compute_precision
            p_gpu.append(chk.precision(int(s), nodes[i], vals[i], en, ev)[0])
        else:
            (O.fora_topk_new if opt else O.fora_topk_with_bound)(int(s))
            on, ov = O.topk_ppr(k)
            p_ref.append(O.precision(k, on, ov, en, ev)[0])
            p_gpu.append(O.precision(k, nodes[i], vals[i], en, ev)[0])
    mg, mr = float(np.mean(p_gpu)), float(np.mean(p_ref))
    assert mg > 0.95 and abs(mg - mr) <= TOPK_TOL, (mg, mr)
