"""Where does the end-to-end path lose time against the device-resident one?  (development aid)
Runs fora_query_batch with pinned host output for several batch sizes and prints the engine's timing."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import fora_b200 as fb

n, m, desc, op, oc, _, _ = bench.make_graph("lj")
q = bench.query_list(n)
E = fb.Engine(0, seed=2026, slots=32)
stream = torch.cuda.current_stream()
E.set_stream(stream.cuda_stream)
E.upload_graph(n, m, op, oc)
E.configure("fora", 0.5, opt=1, balanced=1)
os.system("free -g | head -2; nproc")
# raw D2H bandwidth, pinned
hp = torch.empty(1 << 28, dtype=torch.uint8).pin_memory()
dp = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); hp.copy_(dp, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("pinned D2H %.1f GB/s" % ((1 << 28) / e0.elapsed_time(e1) / 1e6))
NQ = int(sys.argv[1]) if len(sys.argv) > 1 else 192
h_ppr = torch.empty((NQ, n), dtype=torch.float64).pin_memory()
out = h_ppr.numpy()
d_src = torch.from_numpy(q[:NQ].copy()).cuda()
E.query_batch("fora", q[:8], out=out[:8])
for nq in (64, 128, NQ):
    torch.cuda.synchronize(); t = time.time()
    _, st, tm = E.query_batch("fora", q[:nq], out=out[:nq])
    torch.cuda.synchronize(); dt = time.time() - t
    torch.cuda.synchronize(); t = time.time()
    st2, tm2 = E.query_batch_device("fora", d_src.data_ptr(), nq)
    torch.cuda.synchronize(); dt2 = time.time() - t
    print("nq %d: host-out %.1f q/s (%.1f ms; engine total %.1f copy %.1f) | device %.1f q/s (%.1f ms)" % (nq, nq / dt, dt * 1e3, tm["total_ms"], tm["copy_ms"], nq / dt2, dt2 * 1e3))
