"""Time the unmodified reference (oracle/_ref) on the LJ-shape synthetic graph (development script)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fora_b200 as fb
from helpers import Reference
class G: pass
n, m = 4847571, 68993773
if len(sys.argv) > 1 and sys.argv[1] == "ws": n, m = 281904, 2312497
src, dst = fb.synth_edges(n, m, 42)
g = G(); g.n = n; g.m_decl = m
g.out_ptr, g.out_col, g.in_ptr, g.in_col = fb.csr_from_edges(n, src, dst)
opt, bal = (1, 1) if n > 1e6 else (0, 0)
R = Reference(g, epsilon=0.5, opt=opt, balanced=bal)
print(R.setting("fora")); R.init_query_state()
rng = np.random.default_rng(43)
srcs = rng.integers(0, n, 1000).astype(np.int32)
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for s in srcs[:nq]:
    t = time.time(); R.query("fora", int(s)); dt = time.time() - t
    print("source %d time %.2fs total_rw %d push %.2fs walk %.2fs" % (s, dt, R.counters()[0], R.timer(5), R.timer(6)), flush=True)
print("avg query time %.3f" % (R.timer(3) / nq))
