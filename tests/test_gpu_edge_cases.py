"""Edge cases of the CUDA path through the C ABI: empty and ragged inputs, degenerate graphs, maximum slot count,
a hub whose adjacency list is cut across many CTAs."""
import numpy as np
import pytest

import fora_b200 as fb
from helpers import Graph, Oracle

pytestmark = pytest.mark.gpu


def relerr(a, b):
    d = np.abs(a - b)
    s = np.maximum(np.abs(a), np.abs(b))
    m = s > 0
    return float((d[m] / s[m]).max()) if m.any() else 0.0


def test_tiny_and_edgeless_graphs():
    E = fb.Engine(0, seed=1, slots=4)
    # 5 nodes, a cycle plus one dangling node
    src = np.array([0, 1, 2, 3, 0], np.int32)
    dst = np.array([1, 2, 3, 0, 4], np.int32)
    g = Graph(5, src, dst)
    E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col, g.in_ptr, g.in_col)
    rmax, omega = E.configure("fora", 0.5)
    ppr, st, _ = E.query_batch("fora", np.arange(5, dtype=np.int32))
    O = Oracle(g)
    for s in range(5):
        exact = O.power_iteration(s, 300)
        assert abs(ppr[s].sum() - 1) < 1e-9 and np.abs(ppr[s] - exact).max() < 0.05
    assert ppr[4][4] == 1.0
    nodes, vals, iters, _, _ = E.topk_batch("fora", np.array([0], np.int32), 2)
    assert nodes[0][0] == 0 and vals[0][0] > vals[0][1] > 0
    with pytest.raises(fb.ForaError):
        E.topk_batch("fora", np.array([0], np.int32), 4)  # k must be < n-1
    # no edges at all: every node is dangling, ppr = indicator of the source
    g0 = Graph(40, np.empty(0, np.int32), np.empty(0, np.int32), m_decl=1)
    E.upload_graph(g0.n, 1, g0.out_ptr, g0.out_col)
    E.configure("fora", 0.5)
    ppr, st, _ = E.query_batch("fora", np.array([3, 39], np.int32))
    assert ppr[0][3] == 1.0 and ppr[1][39] == 1.0 and ppr.sum() == 2.0 and st[0]["n_walks"] == 0
    d, hops = E.random_walks(7, 100, 0)
    assert (d == 7).all() and hops == 0
    # empty batch
    ppr, st, tm = E.query_batch("fora", np.empty(0, np.int32))
    assert ppr.shape == (0, 40) and st == []
    E.close()


def test_hub_cut_across_ctas_and_dangling_fanout():
    # star: hub 0 -> 100000 leaves (half of them dangling, half pointing back), so one frontier vertex owns
    # 100k edges (cut across many tiles / CTAs) and 50k dangling vertices return their mass to the source
    n = 100001
    leaves = np.arange(1, n, dtype=np.int32)
    back = leaves[::2]
    src = np.concatenate([np.zeros(n - 1, np.int32), back])
    dst = np.concatenate([leaves, np.zeros(len(back), np.int32)])
    g = Graph(n, src, dst)
    E = fb.Engine(0, seed=2, slots=2)
    E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col)
    rmax, omega = E.configure("fora", 0.5)
    O = Oracle(g)
    O.init_state(-1.0, 0)
    for s in (0, 1, 2):
        res, rsd, rsum, st = E.push_only(s, rmax)
        O.reset_counters()
        r2 = O.push_sync(s, rmax, 1, 0)
        a, b = O.fwd()
        c = O.counters()
        # rsum: ours is the sum of the residues, the reference's is 1 - alpha*sum(pushed), which drifts by ~1e-11 over 1e5 subtractions
        assert relerr(res, a) < 1e-9 and relerr(rsd, b) < 1e-9 and abs(rsum - r2) < 1e-10 and abs(rsum - rsd.sum()) < 1e-13
        assert (st["edges_pushed"], st["vertices_pushed"], st["push_levels"]) == (c["edges_pushed"], c["vertices_pushed"], c["push_levels"])
    ppr, st, _ = E.query_batch("fora", np.array([0, 1, 2], np.int32))
    exact = O.power_iteration(0, 300)
    big = exact >= 1.0 / n
    assert (np.abs(ppr[0][big] - exact[big]) / exact[big]).max() < 0.5
    E.close()


def test_max_slots_and_ragged_last_wave():
    g = Graph.synth(3000, 30000, seed=7)
    srcs = np.random.default_rng(0).integers(0, g.n, 70).astype(np.int32)
    outs = []
    for slots in (64, 7):
        E = fb.Engine(0, seed=9, slots=slots)
        E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col)
        E.configure("fora", 0.5, opt=1, balanced=1)
        ppr, st, tm = E.query_batch("fora", srcs)
        outs.append((ppr, [s["n_walks"] for s in st], [s["push_rounds"] for s in st]))
        E.close()
    assert outs[0][1] == outs[1][1] and outs[0][2] == outs[1][2]
    assert np.allclose(outs[0][0], outs[1][0], rtol=1e-9, atol=1e-15)
    with pytest.raises(fb.ForaError):
        fb.Engine(0, slots=65)
    E = fb.Engine(0)
    with pytest.raises(fb.ForaError):
        E.query_batch("fora", srcs)  # no graph / params
    E.close()


@pytest.mark.parametrize("env", [
    {"FORA_PUSH_PACK": "0"},                        # plain columns, degree loaded per edge
    {"FORA_PACK_SHIFT": "29"},                      # 3 degree bits: every target with d_out >= 7 takes the saturated path
    {"FORA_PACK_SHIFT": "31"},                      # 1 degree bit: only d_out = 0 is ever packed
    {"FORA_RELABEL_KEY": "0"},                      # relabelling by plain in-degree
    {"FORA_NO_RELABEL": "1"},                       # no relabelling at all
    {"FORA_WALK_HOT_MB": "0.01"},                   # evict_first instantiation of the walk kernel on a small graph
    {"FORA_WALK_HOT_MB": "0"},                      # unhinted instantiation
    {"FORA_PUSH_LOG": "0"},                         # reserve credited directly in phase A, no log
    {"FORA_PUSH_LOG_CAP": "37"},                    # a log of 37 entries per slot: almost every push overflows into the direct update
    {"FORA_PUSH_LOG_CAP": "2500"},                  # the log fills up in the middle of a push
])
def test_layout_variants_give_the_same_answers(env, monkeypatch):
    """Packed columns, relabelling keys and L2 hints are layout choices: push results must match the schedule-matched
    oracle under each of them, walk counts must not move, and a query must still meet FORA's guarantee."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    g = Graph.synth(6000, 90000, seed=11)
    E = fb.Engine(0, seed=5, slots=3)
    E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col)
    op, oc, _, _ = E.download_csr(with_in=False)
    assert (op == g.out_ptr).all() and (oc == g.out_col).all()
    rmax, omega = E.configure("fora", 0.5, opt=1, balanced=0)
    O = Oracle(g)
    O.init_state(-1.0, 0)
    hub = int(np.argmax(g.deg))
    for s in (hub, 17):
        res, rsd, rsum, st = E.push_only(s, rmax)
        r2 = O.push_sync(s, rmax, 1, 0)
        a, b = O.fwd()
        assert relerr(res, a) < 1e-9 and relerr(rsd, b) < 1e-9 and abs(rsum - r2) < 1e-12
    for balanced in (0, 1):  # balanced: several push rounds per wave, the credit log is applied once after the last one
        E.configure("fora", 0.5, opt=1, balanced=balanced)
        ppr, st, _ = E.query_batch("fora", np.array([hub, 17, 4000], np.int32))
        for i, s in enumerate((hub, 17, 4000)):
            exact = O.power_iteration(int(s), 150)
            big = exact >= 1.0 / g.n
            assert abs(ppr[i].sum() - 1.0) < 1e-9
            assert (np.abs(ppr[i][big] - exact[big]) / exact[big]).max() < 0.5
    E.close()
