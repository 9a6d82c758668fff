"""Per-level trace of one wave of the shipped push kernel with dense slot-levels / edge lists (development script).
usage: FORA_PUSH_DENSE=0.03 python scripts/p1_trace.py [slots]"""
import os, sys, ctypes as C
os.environ["FORA_PUSH_TRACE"] = "1"
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
E.configure("fora", 0.5, opt=1, balanced=0)
srcs = np.random.default_rng(43).integers(0, n, 1000).astype(np.int32)
_, stats, tm = E.query_batch("fora", srcs[:S], want_ppr=False)
ed = sum(s["edges_pushed"] for s in stats)
print("edges %d, push kernel %.3f ms/q, %.2f G edges/s" % (ed, tm["push_kernel_ms"] / S, ed / tm["push_kernel_ms"] / 1e6))
out = np.zeros(4 * 4096, np.uint64)
E.L.fora_debug_push_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
lv = E.L.fora_debug_push_trace(E.h, out.ctypes.data, 4096)
t = out[: 4 * lv].reshape(lv, 4).astype(np.int64)
tx = out[4 * 1024: 4 * 1024 + 4 * lv].reshape(lv, 4).astype(np.int64)
ts = out[4 * 2048: 4 * 2048 + 8 * lv].reshape(lv, 8).astype(np.int64)
for i in range(lv - 1):
    dt = t[i + 1, 0] - t[i, 0]
    nf, nd = int(t[i, 1]) & ((1 << 40) - 1), int(t[i, 1]) >> 40
    line = "L%3d nf(A/B)=%9d E(A/B)=%10d dense %2d  level %9.1f us  phase A %7.1f us" % (i, nf, t[i, 2], nd, dt / 1e3, (t[i, 3] - t[i, 0]) / 1e3)
    if nd:
        line += "   per dense slot, CTA 0: adds %.1f wait %.1f scan+gather %.1f wait %.1f us" % tuple(tx[i, k] / 1e3 / nd for k in range(4))
        line += " | scan stages: pass1 %.1f count %.1f compact %.1f pass2 %.1f degscan %.1f expand %.1f" % tuple(ts[i, k] / 1e3 / nd for k in range(6))
    print(line)
