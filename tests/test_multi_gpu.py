"""One query split over GPUs (fora_b200/multi.py, BASELINE.json config 5): the per-part walk ranges must add up
to exactly the single-GPU result (same Philox keys, same plan), with or without a process group."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import fora_b200 as fb
from helpers import ROOT, Graph, Oracle

pytestmark = pytest.mark.gpu


def test_parts_add_up_to_the_single_gpu_result():
    import torch
    from fora_b200 import multi
    g = Graph.synth(20000, 200000, seed=3)
    s = 11
    outs = {}
    for nparts in (1, 3):
        acc = np.zeros(g.n)
        for part in range(nparts):
            E = fb.Engine(0, seed=5, slots=1)
            E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col)
            rmax, omega = E.configure("fora", 0.5, opt=1)
            res, rsd, rsum, _ = E.push_only(s, rmax)   # state stays in slot 0
            E.compute_ppr_part_device(rsum, 0, part, nparts)
            acc += multi.to_original(E, multi.device_tensor(E.device_reserve_ptr(0), g.n, torch.device("cuda", 0))).cpu().numpy()
            E.close()
        outs[nparts] = acc
    assert np.allclose(outs[1], outs[3], rtol=1e-12, atol=1e-18) and abs(outs[3].sum() - 1.0) < 1e-9
    E = fb.Engine(0, seed=5, slots=1)
    E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col)
    rmax, omega = E.configure("fora", 0.5, opt=1)
    ppr, st = multi.ssppr_split(E, s, rmax)          # no process group: world size 1
    single, _, _ = E.query_batch("fora", np.array([s], np.int32))
    assert np.allclose(ppr.cpu().numpy(), outs[1], rtol=1e-12, atol=1e-18)
    exact = Oracle(g).power_iteration(s, 150)
    big = exact >= 1.0 / g.n
    assert (np.abs(outs[3][big] - exact[big]) / exact[big]).max() < 0.5
    E.close()


def test_nccl_world_size_2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "scripts", "multi_gpu_check.py"), "small"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    import json
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["world"] == 2 and abs(d["sum"] - 1.0) < 1e-9 and d["max_abs_diff_vs_single_gpu"] < 1e-9 and d["max_rel_err_vs_exact"] < 0.5


def test_group_query_split_in_process():
    # fora_group_*: the C++ host path of config 5 -- push on GPU 0, ncclBroadcast of the compacted (vertex, residue) list, walks
    # split by chunk range, ncclAllReduce of the dense vectors, all issued by the library.  A group of ONE GPU runs the same code
    # without collectives; with >= 2 GPUs the result must equal the single-GPU vector up to fp64 summation order.
    import torch
    g = Graph.synth(20000, 200000, seed=3)
    srcs = [11, int(np.argmax(g.deg)), int(np.flatnonzero(g.deg == 0)[0])]
    G1 = fb.Group(1, seed=5)
    G1.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col)
    G1.configure("fora", 0.5, opt=1)
    E = fb.Engine(0, seed=5, slots=1)
    E.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col)
    E.configure("fora", 0.5, opt=1)
    O = Oracle(g)
    base = {}
    for i, s in enumerate(srcs):
        ppr, st, tm = G1.query_split(s, query_id=i)
        E.set_query_base(i)
        single, st1, _ = E.query_batch("fora", np.array([s], np.int32))
        assert np.allclose(ppr, single[0], rtol=1e-12, atol=1e-18) and st["n_walks"] == st1[0]["n_walks"] and st["walk_hops"] == st1[0]["walk_hops"]
        assert abs(ppr.sum() - 1.0) < 1e-9 and tm["n_gpus"] == 1 and tm["reduce_bytes"] == 0
        base[s] = ppr
    G1.close()
    E.close()
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("the multi-GPU half needs 2 GPUs")
    for n_gpus in sorted({2, min(ng, 8)}):
        G = fb.Group(n_gpus, seed=5)
        G.upload_graph(g.n, g.m_decl, g.out_ptr, g.out_col)
        for balanced in (0, 1):
            G.configure("fora", 0.5, opt=1, balanced=balanced)
            for i, s in enumerate(srcs):
                ppr, st, tm = G.query_split(s, query_id=i)
                assert abs(ppr.sum() - 1.0) < 1e-9 and tm["n_gpus"] == n_gpus
                if not balanced:
                    assert np.allclose(ppr, base[s], rtol=1e-12, atol=1e-18)  # same plan, same Philox keys: only the summation order differs
                    if g.deg[s]:
                        assert tm["bcast_bytes"] == 12 * st["n_sources"] and tm["reduce_bytes"] == 8 * g.n
                exact = O.power_iteration(int(s), 150)
                big = exact >= 1.0 / g.n
                assert (np.abs(ppr[big] - exact[big]) / exact[big]).max() < 0.5
        G.close()
