"""Round-2 push experiments: one line per variant (environment knobs read by libfora_b200.so) on the LJ-shape workload.

usage: python scripts/r2_push_sweep.py [--slots S] [--nq N] [--trace] name='{"FORA_X": 1, ...}' ...
FORA eps=0.5 --balanced --opt, fresh engine per variant, `slots` warm-up queries, then N timed queries left on the device.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
args = sys.argv[1:]
SLOTS, NQ, TRACE, BAL = 48, 96, False, 1
while args and args[0].startswith("--"):
    o = args.pop(0)
    if o == "--slots": SLOTS = int(args.pop(0))
    elif o == "--nq": NQ = int(args.pop(0))
    elif o == "--bal": BAL = int(args.pop(0))
    elif o == "--trace": TRACE = True
if TRACE:
    os.environ["FORA_PUSH_TRACE"] = "1"
import fora_b200 as fb  # noqa: E402

if os.environ.get("FORA_VARIANT_LIB"):  # a build with other compile-time constants (fora_b200/variants/lib_<name>.so)
    fb.LIB_PATH = os.path.join(ROOT, "fora_b200", "variants", "lib_%s.so" % os.environ["FORA_VARIANT_LIB"])
    print("library:", fb.LIB_PATH, flush=True)

n, m = 4847571, 68993773
src, dst = fb.synth_edges(n, m, 42)
op, oc, _, _ = fb.csr_from_edges(n, src, dst, with_in=False)
del src, dst
queries = np.random.default_rng(43).integers(0, n, 1000).astype(np.int32)
ref = None
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "r2_push_sweep.jsonl"), "a")
seen = set()
for spec in args or ["default={}"]:
    name, js = spec.split("=", 1)
    env = json.loads(js)
    slots = int(env.pop("slots", SLOTS))
    for k in list(seen):
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ[k] = str(v)
        seen.add(k)
    E = fb.Engine(0, seed=2026, slots=slots)
    try:
        E.upload_graph(n, m, op, oc)
        rmax, _ = E.configure("fora", 0.5, opt=1, balanced=BAL)
        E.query_batch("fora", queries[:slots], want_ppr=False)
        _, stats, tm = E.query_batch("fora", queries[slots:slots + NQ], want_ppr=False)
        ed = sum(s["edges_pushed"] for s in stats) / NQ
        lv = sum(s["push_levels"] for s in stats) / NQ
        hops = sum(s["walk_hops"] for s in stats) / NQ
        r = {"name": name, "env": env, "slots": slots, "nq": NQ, "qps": NQ / (tm["total_ms"] * 1e-3), "push_kernel_ms_per_q": tm["push_kernel_ms"] / NQ,
             "push_phase_ms_per_q": tm["push_ms"] / NQ, "walk_kernel_ms_per_q": tm["walk_kernel_ms"] / NQ, "edges_per_q": ed, "levels_per_q": lv,
             "G_edges_per_s": ed / (tm["push_kernel_ms"] / NQ) / 1e6, "G_hops_per_s": hops / (tm["walk_kernel_ms"] / NQ) / 1e6, "launches": tm["kernel_launches"]}
        # answers: push state of one query (1e-12: atomic order only) and the walk counts of the batch
        res, rsd, rsum, st = E.push_only(int(queries[0]), rmax * 4)
        sig = (res, rsd, [s["n_walks"] for s in stats[:8]], [s["edges_pushed"] for s in stats[:8]], [s["push_levels"] for s in stats[:8]])
        if ref is None:
            ref = sig
            chk = "reference answers"
        else:
            chk = "d reserve %.1e d residue %.1e walks %s edges %s levels %s" % (np.abs(res - ref[0]).max(), np.abs(rsd - ref[1]).max(), sig[2] == ref[2], sig[3] == ref[3], sig[4] == ref[4])
        print("%-26s slots %2d: %6.1f q/s | push kernels %.3f ms/q (%.1f G edges/s, %.1fM edges, %.0f levels) push phase %.3f | walk %.3f ms/q (%.1f G hops/s) | launches %d | %s" % (
            name, slots, r["qps"], r["push_kernel_ms_per_q"], r["G_edges_per_s"], ed / 1e6, lv, r["push_phase_ms_per_q"], r["walk_kernel_ms_per_q"], r["G_hops_per_s"], r["launches"], chk), flush=True)
        LOG.write(json.dumps(r) + "\n")
        LOG.flush()
        if TRACE:
            E.configure("fora", 0.5, opt=1, balanced=0)
            E.query_batch("fora", queries[:slots], want_ppr=False)
            out = np.zeros(4 * 4096, np.uint64)
            E.L.fora_debug_push_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
            lvn = E.L.fora_debug_push_trace(E.h, out.ctypes.data, 4096)
            t = out[: 4 * lvn].reshape(lvn, 4).astype(np.int64)
            for i in range(lvn - 1):
                dt = t[i + 1, 0] - t[i, 0]
                print("  L%3d nf=%8d  level %8.1f us  phaseA+barrier %7.1f us  CTA0 phase B %7.1f us  wait at level barrier %7.1f us" % (
                    i, t[i, 1], dt / 1e3, (t[i, 3] - t[i, 0]) / 1e3, (t[i, 2] - t[i, 3]) / 1e3, (t[i + 1, 0] - t[i, 2]) / 1e3))
    finally:
        E.close()
