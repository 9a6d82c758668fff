"""torchrun --nproc-per-node N scripts/multi_gpu_check.py : split one query's walks over N GPUs, NCCL-reduce the
PPR vector, compare with the single-GPU result and with power iteration (development / profiles)."""
import os, sys, time, json
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fora_b200 as fb
from fora_b200 import multi
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
shape = sys.argv[1] if len(sys.argv) > 1 else "lj"
n, m = {"lj": (4847571, 68993773), "small": (200000, 3000000), "pokec": (1632803, 30622564), "twitter": (41652230, 1468365182)}[shape]
t0 = time.time()
src, dst = fb.synth_edges(n, m, 42)
t1 = time.time()
E = fb.Engine(local, seed=99, slots=1)
E.set_stream(torch.cuda.current_stream().cuda_stream)
if shape == "twitter":
    E.build_graph_from_edges(n, m, src, dst, with_in=False)   # K0 on the device: a host counting sort of 1.5e9 edges takes minutes
    op = E.download_csr(with_in=False)[0]
    oc = None
else:
    op, oc, _, _ = fb.csr_from_edges(n, src, dst, with_in=False)
    E.upload_graph(n, m, op, oc)
del src, dst
if rank == 0: print("graph: synth %.1fs, csr+upload %.1fs" % (t1 - t0, time.time() - t1), flush=True)
rmax, omega = E.configure("fora", 0.5, opt=1)
deg = np.diff(op)
s = int(np.flatnonzero(deg > 5)[4321])
for rep in range(3):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ppr, st = multi.ssppr_split(E, s, rmax, qid=0)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
total = ppr.sum().item()
if rank == 0:
    split = ppr.cpu().numpy().copy()
    E1 = E if shape == "twitter" else fb.Engine(local, seed=99, slots=1)
    if shape != "twitter":
        E1.upload_graph(n, m, op, oc); E1.configure("fora", 0.5, opt=1)
    single, st1, _ = E1.query_batch("fora", np.array([s], np.int32))
    exact = E1.power_iteration(s, 100)
    big = exact >= 1.0 / n
    print(json.dumps({"world": world, "shape": shape, "seconds_per_query": dt, "sum": total, "max_abs_diff_vs_single_gpu": float(np.abs(split - single[0]).max()),
                      "max_rel_err_vs_exact": float((np.abs(split[big] - exact[big]) / exact[big]).max()), "walks_single": st1[0]["n_walks"], "walks_plan": st["n_walks"], "rmax": rmax, "omega": omega, "n": n, "m": m}))
if world > 1: dist.destroy_process_group()
