"""Per-level trace of one push launch on the LJ-shape graph (development script)."""
import os, sys, ctypes as C
os.environ["FORA_PUSH_TRACE"] = "1"
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fora_b200 as fb
n, m = 4847571, 68993773
src, dst = fb.synth_edges(n, m, 42)
op, oc, _, _ = fb.csr_from_edges(n, src, dst, with_in=False)
E = fb.Engine(0, seed=7, slots=int(sys.argv[1]) if len(sys.argv) > 1 else 1)
E.upload_graph(n, m, op, oc)
rmax, omega = E.configure("fora", 0.5, opt=1, balanced=0)
rng = np.random.default_rng(43)
srcs = rng.integers(0, n, 1000).astype(np.int32)
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1
_, stats, tm = E.query_batch("fora", srcs[:nq], want_ppr=False)
_, stats, tm = E.query_batch("fora", srcs[:nq], want_ppr=False)
print(tm, stats[0])
if len(sys.argv) > 3: sys.exit(0)
out = np.zeros(4 * 4096, np.uint64)
E.L.fora_debug_push_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
lv = E.L.fora_debug_push_trace(E.h, out.ctypes.data, 4096)
t = out[: 4 * lv].reshape(lv, 4).astype(np.int64)
print("levels", lv)
tot = 0
for i in range(lv):
    dt = (t[i + 1, 0] - t[i, 0]) if i + 1 < lv else 0
    da = t[i, 3] - t[i, 0]
    tot += dt
    print("L%3d nf=%8d E=%9d  level %8.1f us  phaseA %7.1f us  -> %6.2f G edges/s" % (i, t[i, 1], t[i, 2], dt / 1e3, da / 1e3, t[i, 2] / max(dt, 1)))
print("sum %.1f us" % (tot / 1e3))
A = sum((t[i, 3] - t[i, 0]) for i in range(lv - 1)) / 1e3
L = sum((t[i + 1, 0] - t[i, 0]) for i in range(lv - 1)) / 1e3
small = sum((t[i + 1, 0] - t[i, 0]) for i in range(lv - 1) if t[i, 2] < 2000000) / 1e3
print("phase A %.1f us (%.0f %%), phase B + barriers %.1f us, levels with < 2M edges: %.1f us (%.0f %%), entries %d edges %d" % (
    A, 100 * A / L, L - A, small, 100 * small / L, t[:, 1].sum(), t[:, 2].sum()))
