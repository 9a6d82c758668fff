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
