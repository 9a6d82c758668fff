"""Experiment: two engine contexts ("lanes") on ONE GPU, so that the push phase of one wave overlaps the walk
phase of another.  usage: lanes_experiment.py <lanes> <slots> <queries per lane>"""
import os, sys, time, threading
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import fora_b200 as fb

lanes, slots, nq = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n, m, desc, op, oc, _, _ = bench.make_graph("lj")
q = bench.query_list(n)
engs = []
for l in range(lanes):
    E = fb.Engine(0, seed=2026, slots=slots)
    E.upload_graph(n, m, op, oc)
    E.configure("fora", 0.5, opt=1, balanced=1)
    engs.append(E)
d_src = [torch.from_numpy(q[l * nq:(l + 1) * nq].copy()).cuda() for l in range(lanes)]
res = [None] * lanes
def work(l, reps):
    for _ in range(reps):
        res[l] = engs[l].query_batch_device("fora", d_src[l].data_ptr(), nq)
for reps in (1, 3):
    torch.cuda.synchronize(); t = time.time()
    th = [threading.Thread(target=work, args=(l, reps)) for l in range(lanes)]
    [x.start() for x in th]; [x.join() for x in th]
    torch.cuda.synchronize(); dt = time.time() - t
    tm = res[0][1]
    print("lanes %d slots %d: %d queries in %.1f ms -> %.1f q/s   (lane0 push %.1f ms walk %.1f ms, kernels push %.1f walk %.1f)" % (
        lanes, slots, lanes * nq * reps, dt * 1e3, lanes * nq * reps / dt, tm["push_ms"], tm["walk_ms"], tm["push_kernel_ms"], tm["walk_kernel_ms"]))
