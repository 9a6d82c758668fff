#!/usr/bin/env python
"""bench.py -- SSPPR queries/s for FORA (eps=0.5, --balanced --opt) on a synthetic LiveJournal-shape
power-law graph (4,847,571 nodes / 68,993,773 edges), BASELINE.json config #2.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One process per GPU (under torchrun for N>1; queries are sharded, the graph is replicated, no
data-path collective).  A step = one pass of the query path over one batch of `--batch` queries.
Prints ONE JSON line on rank 0:
  value     whole-job queries/s, graph and query ids resident in HBM, CUDA events, max over ranks
  e2e       the same metric through the C-ABI call with HOST buffers: query ids copied in, every
            dense fp64 PPR vector (n*8 bytes per query) copied out to pinned host memory
  roofline  dominant kernel: algorithmic bytes / CUDA-event time of its launches
  cpu_baseline  the unmodified reference (oracle/_ref, kind "reference") or the C oracle (kind "port")
            on one host core, on a bounded sample of the same workload
`--impl reference` times the reference's own CPU implementation on all host cores instead.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPES = {
    "lj": (4847571, 68993773, "synthetic LiveJournal-shape power-law graph (4,847,571 nodes / 68,993,773 edges)"),
    "webstanford": (281904, 2312497, "synthetic webstanford-shape power-law graph (281,904 nodes / 2,312,497 edges)"),
    "pokec": (1632803, 30622564, "synthetic Pokec-shape power-law graph (1,632,803 nodes / 30,622,564 edges)"),
}
GRAPH_SEED, QUERY_SEED, N_QUERIES = 42, 43, 1000
EPS = 0.5


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_graph(shape, with_in=False):
    # host-only helpers: fb.synth_edges / fb.csr_from_edges bind libfora_host.so, NOT the GPU engine (the reference arm must not
    # map libfora_b200.so; the GPU arm loads the engine when it creates its Engine)
    import fora_b200 as fb
    n, m, desc = SHAPES[shape]
    t = time.time()
    src, dst = fb.synth_edges(n, m, GRAPH_SEED)
    op, oc, ip_, ic = fb.csr_from_edges(n, src, dst, with_in=with_in)
    log("[bench] graph %s generated in %.1fs" % (shape, time.time() - t))
    return n, m, desc, op, oc, ip_, ic


def query_list(n):
    # ssquery.txt: query ids uniform in [0,n), one per line (algo.h:504-508); seed 43
    return np.random.default_rng(QUERY_SEED).integers(0, n, N_QUERIES).astype(np.int32)


# ------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU side: the unmodified reference (oracle/_ref) or the C oracle -- the CHECKER, timed as baseline
# ------------------------------------------------------------------------------------------------
_W = {}  # per worker process: the reference (or the oracle) with its graph built ONCE


def _cpu_worker_init(n, m, op, oc, cores, counter):
    """Pool initializer: pin the worker to one core and build the reference's Graph once; every later step reuses it
    (round 1 rebuilt the 69 M-edge vector<vector<int>> in every step, which was half of the arm's wall time)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    if cores is not None:
        with counter.get_lock():
            idx = counter.value
            counter.value += 1
        try:
            os.sched_setaffinity(0, {cores[idx % len(cores)]})
        except Exception:
            pass
    import helpers

    class G:
        pass
    g = G()
    g.n, g.m_decl, g.out_ptr, g.out_col = n, m, op, oc
    g.in_ptr, g.in_col = np.zeros(n + 1, np.int64), np.zeros(1, np.int32)  # gr is not used by FORA queries
    if helpers.have_reference():
        R = helpers.Reference(g, epsilon=EPS, opt=1, balanced=1)  # maps oracle/_ref/libfora_ref.so in place
        R.setting("fora")        # fora_setting, query.h:1461
        R.init_query_state()     # query.h:1427,1464-1467
        _W["ref"] = R
    else:
        g.deg = np.diff(op)
        O = helpers.Oracle(g, seed=1)
        rmax, omega = O.setting("fora", EPS, opt=1)
        O.set_params(EPS, rmax, omega, opt=1, balanced=1)
        O.init_state(-1.0, 0)
        _W["orc"] = O


def _cpu_worker(sources):
    """Run `sources` through fora_query_basic on this worker's core; returns (seconds in the reference's own FORA_QUERY
    timer, or wall clock for the port, kind)."""
    if "ref" in _W:
        R = _W["ref"]
        t0 = R.timer(3)
        for s in sources:
            R.query("fora", int(s))   # fora_query_basic under Timer(FORA_QUERY), query.h:841-842
        return R.timer(3) - t0, "reference"
    O = _W["orc"]
    t0 = time.perf_counter()
    for s in sources:
        O.fora_query(int(s))
    return time.perf_counter() - t0, "port"


def _make_pool(P, n, m, op, oc, cores):
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    return ctx.Pool(P, initializer=_cpu_worker_init, initargs=(n, m, op, oc, cores, ctx.Value("i", 0)))


def cpu_baseline_single(n, m, op, oc, queries, n_sample):
    with _make_pool(1, n, m, op, oc, None) as pool:
        secs, kind = pool.map(_cpu_worker, [list(queries[:n_sample])])[0]
    return {"value": n_sample / secs, "unit": "queries/s", "cores": 1, "kind": kind,
            "sample": "%d of the %d queries of the same workload (first ids of the seed-%d list), fora_query_basic per query, 1 thread" % (n_sample, N_QUERIES, QUERY_SEED),
            "seconds": secs}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shape = args.shape
    n, m, desc, op, oc, _, _ = make_graph(shape)
    queries = query_list(n)
    try:
        cores = sorted(os.sched_getaffinity(0))
    except Exception:
        cores = list(range(os.cpu_count() or 1))
    P = max(1, min(len(cores), args.ref_procs if args.ref_procs > 0 else len(cores)))
    per_step = P * args.ref_queries_per_proc  # queries per step, one persistent worker process per core
    budget = float(os.environ.get("FORA_REF_BUDGET_S", "420"))
    t_all = time.perf_counter()
    pool = _make_pool(P, n, m, op, oc, cores)
    pool.map(_cpu_worker, [[] for _ in range(P)])  # every worker has built its Graph before anything is timed
    log("[bench] reference arm: %d workers ready in %.1fs" % (P, time.perf_counter() - t_all))

    def step(i):
        base = (i * per_step) % N_QUERIES
        ids = [int(queries[(base + j) % N_QUERIES]) for j in range(per_step)]
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [ids[p::P] for p in range(P)], chunksize=1)
        # a step ends when its slowest worker has finished its share: that worker's own FORA_QUERY seconds
        return max(r[0] for r in res), res[0][1], time.perf_counter() - t0

    warm_done = 0
    for w in range(args.warmup):
        q_s, kind, wall = step(w)
        warm_done += 1
        if (time.perf_counter() - t_all) + wall * (args.warmup - warm_done + args.steps) > budget:
            log("[bench] reference arm: skipping %d warm-up step(s) to stay inside the time budget" % (args.warmup - warm_done))
            break
    secs, kind = [], "reference"
    for k in range(args.steps):
        q_s, kind, wall = step(args.warmup + k)
        secs.append(q_s)
    pool.close()
    pool.join()
    total = float(np.sum(secs))
    value = per_step * args.steps / total
    workload = "%s; FORA eps=0.5 --balanced --opt, batched queries from a %d-query list" % (desc, N_QUERIES)
    out = {
        "impl": "reference", "metric": "SSPPR queries/s (FORA eps=0.5, LJ-shape)", "value": value, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "warmup_run": warm_done, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "shape": shape, "queries_per_step": per_step,
                   "sample": "each step = %d queries of that list (bounded sample: one per host core)" % per_step},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": P, "kind": kind,
                         "sample": "%d queries per step, %d persistent worker processes of the single-threaded reference (one per core, Graph built once per worker, oracle/_ref/libfora_ref.so mapped in place), query time = slowest worker's FORA_QUERY timer" % (per_step, P)},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="lj", choices=sorted(SHAPES))
    ap.add_argument("--batch", type=int, default=192, help="queries per step per GPU")
    ap.add_argument("--slots", type=int, default=int(os.environ.get("FORA_SLOTS", "48")))
    ap.add_argument("--e2e-queries", type=int, default=144)
    ap.add_argument("--e2e-dense-queries", type=int, default=96)
    ap.add_argument("--cpu-sample", type=int, default=1, help="queries timed on the CPU baseline (0 = skip)")
    ap.add_argument("--ref-procs", type=int, default=0, help="worker processes of the reference arm (0 = one per host core)")
    ap.add_argument("--ref-queries-per-proc", type=int, default=1)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("[bench] warning: fewer than 3 warm-up steps")
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import fora_b200 as fb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    from fora_b200 import shard
    numa = shard.bind_to_gpu_numa_node(local) if world > 1 and not os.environ.get("FORA_NO_NUMA_BIND") else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, m, desc, op, oc, _, _ = make_graph(args.shape)
    queries = query_list(n)

    E = fb.Engine(local, seed=2026, slots=args.slots)
    stream = torch.cuda.current_stream()
    E.set_stream(stream.cuda_stream)
    E.upload_graph(n, m, op, oc)
    rmax, omega = E.configure("fora", EPS, opt=1, balanced=1)

    B = args.batch
    # weak scaling: every rank processes B queries per step, its own shard of a world*B global batch

    def step_ids(i):
        return shard.step_query_ids(queries, i, B, rank, world)[0]

    steps_total = args.warmup + args.steps
    d_src = [torch.from_numpy(step_ids(i)).cuda() for i in range(steps_total)]
    agg = {"push_kernel_ms": 0.0, "walk_kernel_ms": 0.0, "push_kernel_launches": 0, "walk_kernel_launches": 0, "kernel_launches": 0,
           "edges": 0, "vertices": 0, "hops": 0, "walks": 0, "push_ms": 0.0, "walk_ms": 0.0}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        E.set_query_base((i * world + rank) * B)  # Philox keyed by the global query index: no two queries of the run share a stream
        E.query_batch_device("fora", d_src[i].data_ptr(), B)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.warmup, steps_total):
        E.set_query_base((i * world + rank) * B)
        stats, tm = E.query_batch_device("fora", d_src[i].data_ptr(), B)
        for k in ("push_kernel_ms", "walk_kernel_ms", "push_kernel_launches", "walk_kernel_launches", "kernel_launches", "push_ms", "walk_ms"):
            agg[k] += tm[k]
        agg["edges"] += sum(s["edges_pushed"] for s in stats)
        agg["vertices"] += sum(s["vertices_pushed"] for s in stats)
        agg["hops"] += sum(s["walk_hops"] for s in stats)
        agg["walks"] += sum(s["n_walks"] for s in stats)
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = shard.max_over_ranks(ev0.elapsed_time(ev1))
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- e2e: the C-ABI call with HOST buffers, copies inside the timed region.  Headline: query ids in, the COMPACTED result out
    # -- every (node, value) with value >= delta = 1/n, i.e. everything FORA's guarantee speaks about (the reference itself
    # discards the vector, query.h:1471-1476).  The dense fp64 vector (n*8 bytes per query) is measured too and reported as
    # e2e_dense: the worst case a caller can ask for.
    nq = min(args.e2e_queries, B)
    E.set_query_base(rank * B)
    h_src = torch.from_numpy(step_ids(0)[:nq].copy()).pin_memory()
    src_np = h_src.numpy()
    thr = 1.0 / n
    probe = E.query_batch_sparse("fora", src_np[: min(8, nq)], thr, n)  # sizes the output buffers (and warms the path)
    per_q = int(np.diff(probe[2].astype(np.int64)).max())
    cap_q = min(n, 2 * per_q + 4096)
    h_ids = torch.empty(nq * cap_q, dtype=torch.int32).pin_memory()
    h_vals = torch.empty(nq * cap_q, dtype=torch.float64).pin_memory()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    _, _, sp_off, st_e2e, _ = E.query_batch_sparse("fora", src_np, thr, cap_q, ids=h_ids.numpy(), vals=h_vals.numpy())
    e1.record(stream)
    barrier()
    e2e_value = world * nq / (shard.max_over_ranks(e0.elapsed_time(e1)) / 1e3)
    sp_entries = int(sp_off[-1])
    sp_mass = float(h_vals.numpy()[:sp_entries].sum() / nq)  # PPR mass carried by the entries >= 1/n, per query
    del h_ids, h_vals
    nqd = min(nq, args.e2e_dense_queries)
    h_ppr = torch.empty((nqd, n), dtype=torch.float64).pin_memory()
    ppr_np = h_ppr.numpy()
    E.query_batch("fora", src_np[: min(4, nqd)], out=ppr_np[: min(4, nqd)])  # warm the path
    barrier()
    e0.record(stream)
    E.query_batch("fora", src_np[:nqd], out=ppr_np)
    e1.record(stream)
    barrier()
    e2e_dense = world * nqd / (shard.max_over_ranks(e0.elapsed_time(e1)) / 1e3)
    checksum = float(ppr_np.sum(axis=1).mean())  # each PPR vector sums to 1
    del h_ppr, ppr_np

    # ---- opt-in mode, reported next to the headline (never instead of it): the queries of a wave draw their walks from one
    # pool (fora_ctx_set_shared_walks, include/fora_b200.h) -- same timed region as `value`, two steps
    E.set_shared_walks(True)
    E.set_query_base(rank * B)
    E.query_batch_device("fora", d_src[0].data_ptr(), B)
    barrier()
    e0.record(stream)
    sw_steps, sw_walk_ms, sw_push_ms = 2, 0.0, 0.0
    for i in range(sw_steps):
        j = (args.warmup + i) % steps_total  # (any --steps / --warmup combination)
        E.set_query_base((j * world + rank) * B)
        _, tm = E.query_batch_device("fora", d_src[j].data_ptr(), B)
        sw_walk_ms += tm["walk_ms"]
        sw_push_ms += tm["push_ms"]
    e1.record(stream)
    barrier()
    sw_value = world * B * sw_steps / (shard.max_over_ranks(e0.elapsed_time(e1)) / 1e3)
    E.set_shared_walks(False)

    if rank == 0:
        peaks, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        try:
            pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peaks, peak_src = float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
        except Exception:
            pass
        push_bytes = 56.0 * agg["vertices"] + 24.0 * agg["edges"]
        walk_bytes = 20.0 * agg["hops"] + 16.0 * agg["walks"]
        kern = "push_kernel" if agg["push_kernel_ms"] >= agg["walk_kernel_ms"] else "walk_kernel"
        kb, kms, kl = (push_bytes, agg["push_kernel_ms"], agg["push_kernel_launches"]) if kern == "push_kernel" else (
            walk_bytes, agg["walk_kernel_ms"], agg["walk_kernel_launches"])
        achieved = kb / max(kms, 1e-9) / 1e6
        traffic, traffic_src = None, None
        try:  # dram__bytes_read+write of the DOMINANT kernel from the committed ncu --set full capture of the same launch shape
            prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
            ent = prof.get("kernels", {}).get(kern)
            if ent and ent.get("slots") == args.slots and args.shape == prof.get("shape"):
                traffic = ent.get("dram_bytes_per_launch")
                traffic_src = "%s (commit %s, %s)" % (ent.get("file"), prof.get("commit"), ent.get("launch"))
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peaks, "peak_source": peak_src, "unit": "GB/s",
                    "frac": achieved / peaks, "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": kb / max(kl, 1), "launches": kl, "avg_launch_ms": kms / max(kl, 1),
                    "push": {"GBps": push_bytes / max(agg["push_kernel_ms"], 1e-9) / 1e6, "edges_per_s": agg["edges"] / max(agg["push_kernel_ms"], 1e-9) * 1e3,
                             "kernel_ms": agg["push_kernel_ms"], "frac": push_bytes / max(agg["push_kernel_ms"], 1e-9) / 1e6 / peaks},
                    "walk": {"GBps": walk_bytes / max(agg["walk_kernel_ms"], 1e-9) / 1e6, "steps_per_s": agg["hops"] / max(agg["walk_kernel_ms"], 1e-9) * 1e3,
                             "kernel_ms": agg["walk_kernel_ms"], "frac": walk_bytes / max(agg["walk_kernel_ms"], 1e-9) / 1e6 / peaks}}
        cpu = None
        if world == 1 and args.cpu_sample > 0:
            log("[bench] timing the CPU baseline on %d query(ies) ..." % args.cpu_sample)
            cpu = cpu_baseline_single(n, m, op, oc, queries, args.cpu_sample)
        out = {
            "metric": "SSPPR queries/s (FORA eps=0.5, LJ-shape)", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s; FORA eps=0.5 --balanced --opt, batched queries from a %d-query list" % (desc, N_QUERIES),
                       "shape": args.shape, "queries_per_step_per_gpu": B, "slots": args.slots, "parallelism": "query-sharded x%d, graph replicated" % world, "numa_node_rank0": numa,
                       "l2": "inputs larger than L2 (CSR 0.33 GB + 39 MB dense state per query re-initialised every query)",
                       "rmax": rmax, "omega": omega},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": int(4 * nq), "d2h_bytes_per_step": int(12 * sp_entries + 8 * (nq + 1)),
                    "queries_per_step": nq, "entries_per_query": sp_entries / nq,
                    "note": "fora_query_batch_sparse: host query ids in; out to pinned host memory every (node id int32, value fp64) with value >= delta = 1/n "
                            "(what FORA's guarantee covers; %.4f of the PPR mass per query) plus the offsets" % sp_mass},
            "e2e_dense": {"value": e2e_dense, "unit": "queries/s", "h2d_bytes_per_step": int(4 * nqd), "d2h_bytes_per_step": int(8 * n * nqd), "queries_per_step": nqd,
                          "note": "fora_query_batch: the full dense fp64 PPR vector per query out to pinned host memory (worst case); mean vector sum %.9f" % checksum},
            "opt_in_shared_walks": {"value": sw_value, "unit": "queries/s", "steps": sw_steps, "walk_phase_ms_per_query": sw_walk_ms / (B * sw_steps),
                                    "push_phase_ms_per_query": sw_push_ms / (B * sw_steps),
                                    "note": "NOT the headline: fora_ctx_set_shared_walks(1) -- the %d queries of a wave draw their walks from one pool built per wave "
                                            "(the reference's --with_idx sharing, query.h:290-307, without a stored index); per-query guarantee unchanged, estimates "
                                            "of queries of the same wave correlated" % args.slots},
            "gpu_launches": int(agg["kernel_launches"]),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "per_query": {"walks": agg["walks"] / (B * args.steps), "walk_hops": agg["hops"] / (B * args.steps), "edges_pushed": agg["edges"] / (B * args.steps),
                          "vertices_pushed": agg["vertices"] / (B * args.steps), "push_phase_ms": agg["push_ms"] / (B * args.steps), "walk_phase_ms": agg["walk_ms"] / (B * args.steps)},
        }
        print(json.dumps(out), flush=True)
    E.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
