"""Per-level trace of one push3 wave on the LJ-shape graph (development script): level, frontier, groups, time, time per interval.
usage: python scripts/p3_trace.py [slots] [balanced]"""
import os, sys, ctypes as C
os.environ["FORA_PUSH_TRACE"] = "1"
os.environ.setdefault("FORA_PUSH_V", "3")
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fora_b200 as fb
n, m = 4847571, 68993773
src, dst = fb.synth_edges(n, m, 42)
op, oc, _, _ = fb.csr_from_edges(n, src, dst, with_in=False)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 48
E = fb.Engine(0, seed=7, slots=S)
E.upload_graph(n, m, op, oc)
rmax, omega = E.configure("fora", 0.5, opt=1, balanced=int(sys.argv[2]) if len(sys.argv) > 2 else 0)
srcs = np.random.default_rng(43).integers(0, n, 1000).astype(np.int32)
_, stats, tm = E.query_batch("fora", srcs[:S], want_ppr=False)
_, stats, tm = E.query_batch("fora", srcs[:S], want_ppr=False)
print({k: round(v, 3) if isinstance(v, float) else v for k, v in tm.items()})
ed = sum(s["edges_pushed"] for s in stats)
print("edges %d, push kernel %.3f ms/q, %.2f G edges/s" % (ed, tm["push_kernel_ms"] / S, ed / tm["push_kernel_ms"] / 1e6))
out = np.zeros(4 * 4096, np.uint64)
E.L.fora_debug_push_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
lv = E.L.fora_debug_push_trace(E.h, out.ctypes.data, 4096)
t = out[: 4 * lv].reshape(lv, 4).astype(np.int64)
tx = out[4 * 1024: 4 * 1024 + 4 * lv].reshape(lv, 4).astype(np.int64)
tot = 0
for i in range(lv - 1):
    dt = t[i + 1, 0] - t[i, 0]
    tot += dt
    ng, nd = int(t[i, 2]) & 0xffffffff, int(t[i, 2]) >> 32
    print("L%3d nf=%9d groups %3d (dense %2d)  level %9.1f us  first interval (A only) %6.1f us  per interval %6.1f us" % (
        i, t[i, 1], ng, nd, dt / 1e3, (t[i, 3] - t[i, 0]) / 1e3, dt / 1e3 / (ng + 1)) + ("   dense intervals, CTA 0: B %.1f wait %.1f scan %.1f us each" % (
            tx[i, 0] / 1e3 / nd, tx[i, 1] / 1e3 / nd, tx[i, 2] / 1e3 / nd) if nd else ""))
print("sum %.1f us over %d levels" % (tot / 1e3, lv))
