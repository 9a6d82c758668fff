"""The CPU oracle against the committed golden vectors (tests/golden/ref_small.npz), which were
produced by the unmodified reference (tests/golden/make_golden.py).  Needs neither the reference nor
a GPU, so it also runs on the GPU box."""
import os

import numpy as np
import pytest

from helpers import GOLDEN_DIR, Graph, Oracle


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN_DIR, "ref_small.npz"))


@pytest.fixture(scope="module")
def g(gold):
    return Graph(int(gold["n"]), gold["src"], gold["dst"], int(gold["m_decl"]))


def test_csr_bit_exact(gold, g):
    assert np.array_equal(g.out_ptr, gold["out_ptr"]) and np.array_equal(g.out_col, gold["out_col"])
    assert np.array_equal(g.in_ptr, gold["in_ptr"]) and np.array_equal(g.in_col, gold["in_col"])
    # the oracle's C restatement of the same construction
    import ctypes as C
    from helpers import _p, c_ip, c_lp
    O = Oracle(g)
    op, ip_ = np.zeros(g.n + 1, np.int64), np.zeros(g.n + 1, np.int64)
    oc, ic = np.zeros(len(g.out_col), np.int32), np.zeros(len(g.out_col), np.int32)
    O.lib.orc_csr_from_edges.argtypes = [C.c_int, C.c_longlong, c_ip, c_ip, c_lp, c_ip, c_lp, c_ip]
    O.lib.orc_csr_from_edges(g.n, len(g.src), _p(g.src, c_ip), _p(g.dst, c_ip), _p(op, c_lp), _p(oc, c_ip), _p(ip_, c_lp), _p(ic, c_ip))
    assert np.array_equal(op, gold["out_ptr"]) and np.array_equal(oc, gold["out_col"])
    assert np.array_equal(ip_, gold["in_ptr"]) and np.array_equal(ic, gold["in_col"])


def test_settings(gold, g):
    O = Oracle(g)
    for opt in (0, 1):
        for w in ("fora", "fora_topk", "montecarlo", "bippr", "fwdpush"):
            ref = gold["setting_%s_opt%d" % (w, opt)]
            rmax, omega = O.setting(w, float(gold["eps"]), opt=opt)
            if w != "montecarlo":
                assert rmax == ref[0]
            if w != "fwdpush":
                assert omega == ref[1]


def test_fifo_push_and_power_iteration(gold, g):
    O = Oracle(g)
    O.init_state(-1.0, 0)
    rmax = float(gold["rmax"])
    for i, s in enumerate(gold["sources"]):
        rs = O.push_fifo(int(s), rmax)
        a, b = O.fwd()
        assert rs == float(gold["fifo_rsum_%d" % i])
        assert np.array_equal(a, gold["fifo_reserve_%d" % i]) and np.array_equal(b, gold["fifo_residue_%d" % i])
        assert np.array_equal(O.residue_occur(), gold["fifo_residue_occur_%d" % i])
        assert np.abs(O.power_iteration(int(s), 100) - gold["power_%d" % i]).max() < 1e-13


def test_resumable_rounds(gold, g):
    O = Oracle(g)
    O.init_state(-1.0, 0)
    rmax = float(gold["rmax"])
    s = int(gold["sources"][2])
    O.push_topk_begin(s)
    for k in range(5):
        rs = O.push_topk_round(s, rmax * 8 / 2 ** k, rmax)
        a, b = O.fwd()
        assert rs == float(gold["round_rsum_%d" % k])
        assert np.array_equal(a, gold["round_reserve_%d" % k]) and np.array_equal(b, gold["round_residue_%d" % k])
        assert np.array_equal(O.push_topk_candidates(), gold["round_cand_%d" % k])


def test_reverse_push_index_info_fwdpush(gold, g):
    O = Oracle(g)
    s = int(gold["sources"][1])
    for j in range(2):
        O.reverse_push(s, float(gold["bwd_rmax_%d" % j]), 1.0, 0)
        a, b = O.bwd()
        assert np.array_equal(a, gold["bwd_reserve_%d" % j]) and np.array_equal(b, gold["bwd_residue_%d" % j])
    for opt in (0, 1):
        rmax, omega = O.setting("fora", float(gold["eps"]), opt=opt)
        O.set_params(float(gold["eps"]), rmax, omega, opt=opt)
        off, cnt, total = O.index_info()
        assert np.array_equal(off, gold["idx_off_opt%d" % opt]) and np.array_equal(cnt, gold["idx_cnt_opt%d" % opt])
    rmax, _ = O.setting("fwdpush", float(gold["eps"]))
    O.set_params(float(gold["eps"]), rmax, 0.0)
    O.init_state(0.0, 0)
    O.fwdpush_query(s)
    assert np.array_equal(O.ppr(), gold["fwdpush_ppr"])
