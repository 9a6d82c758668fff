"""Generates tests/golden/ref_small.npz from the UNMODIFIED reference (oracle/_ref/libfora_ref.so,
built by `make -C oracle ref` where /root/reference exists).  The reference itself cannot travel to
the GPU box and ships no golden vectors of its own, so the vectors it produces here are committed:
everything deterministic the query path computes on one small graph.

    python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import Graph, Reference, write_dataset  # noqa: E402

N, M, SEED, EPS = 400, 3200, 21, 0.5


def main():
    g = Graph.synth(N, M, seed=SEED, self_loops=7)
    out = {"n": N, "src": g.src, "dst": g.dst, "m_decl": g.m_decl, "eps": EPS}
    d = tempfile.mkdtemp()
    write_dataset(d, g.n, g.m_decl, g.src, g.dst)
    R = Reference(folder=d + "/", epsilon=EPS)
    op, oc, ip_, ic = R.graph_dump()
    out.update(out_ptr=op, out_col=oc, in_ptr=ip_, in_col=ic)
    for opt in (0, 1):
        Ro = Reference(g, epsilon=EPS, opt=opt)
        for w in ("fora", "fora_topk", "montecarlo", "bippr", "fwdpush"):
            out["setting_%s_opt%d" % (w, opt)] = np.array(Ro.setting(w))
        Ro.setting("fora")
        off, cnt, dest = Ro.build_index(tempfile.mkdtemp())
        out["idx_off_opt%d" % opt], out["idx_cnt_opt%d" % opt] = off, cnt
    rmax, omega = R.setting("fora")
    R.init_query_state()
    deg = np.diff(op)
    srcs = np.array([0, 3, 11, int(np.argmax(deg)), int(np.flatnonzero(deg == 0)[0])], np.int32)
    out["sources"] = srcs
    out["rmax"], out["omega"] = rmax, omega
    for i, s in enumerate(srcs):
        rs = R.push(int(s), rmax)
        a, b, ao, bo = R.fwd()
        out["fifo_rsum_%d" % i], out["fifo_reserve_%d" % i], out["fifo_residue_%d" % i] = rs, a, b
        out["fifo_residue_occur_%d" % i] = bo
        out["power_%d" % i] = R.power_iteration(int(s))
    # resumable rounds
    s = int(srcs[2])
    R.push_topk_begin(s)
    for k in range(5):
        rs = R.push_topk_round(s, rmax * 8 / 2 ** k, rmax)
        a, b, _, _ = R.fwd()
        out["round_rsum_%d" % k], out["round_reserve_%d" % k], out["round_residue_%d" % k] = rs, a, b
        out["round_cand_%d" % k] = R.push_topk_candidates()
    # backward push
    for j, rm in enumerate([0.3, 1e-3]):
        R.lib.ref_set_rmax_omega(rm, 1000.0)
        R.reverse_push(int(srcs[1]))
        a, b = R.bwd()
        out["bwd_reserve_%d" % j], out["bwd_residue_%d" % j], out["bwd_rmax_%d" % j] = a, b, rm
    # fwdpush query (deterministic)
    Rf = Reference(g, epsilon=EPS)
    Rf.setting("fwdpush")
    Rf.init_query_state()
    Rf.query("fwdpush", int(srcs[1]))
    out["fwdpush_ppr"] = Rf.ppr()
    np.savez_compressed(os.path.join(HERE, "ref_small.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_small.npz"), os.path.getsize(os.path.join(HERE, "ref_small.npz")), "bytes")


if __name__ == "__main__":
    main()
