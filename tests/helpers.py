"""Test helpers: ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when it has been
built, the unmodified reference behind oracle/ref_harness.cpp (oracle/_ref/libfora_ref.so).

TEST INFRASTRUCTURE ONLY -- nothing under fora_b200/ imports this module.
"""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libfora_ref.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "fora")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_lp = C.POINTER(C.c_longlong)
c_up = C.POINTER(C.c_ulonglong)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
        os.path.join(ORACLE_DIR, "fora_oracle.c")
    ):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
    return ORACLE_SO


def have_reference():
    return os.path.exists(REF_SO)


# ----------------------------------------------------------------------------- synthetic graphs
def synth_edges(n, m, seed=42, exponent=2.3, dangling_frac=0.03, self_loops=0):
    """Directed power-law (Chung-Lu style) edge list in 'file order', ids randomly permuted,
    duplicates allowed, a few zero-out-degree nodes, optional self loops (the loader drops them)."""
    rng = np.random.default_rng(seed)
    w_out = (np.arange(1, n + 1, dtype=np.float64)) ** (-1.0 / (exponent - 1.0))
    w_in = w_out.copy()
    rng.shuffle(w_in)
    dang = rng.random(n) < dangling_frac
    w_out = np.where(dang, 0.0, w_out)
    perm = rng.permutation(n)
    cdf_o = np.cumsum(w_out / w_out.sum())
    cdf_i = np.cumsum(w_in / w_in.sum())
    src = np.minimum(np.searchsorted(cdf_o, rng.random(m)), n - 1)
    dst = np.minimum(np.searchsorted(cdf_i, rng.random(m)), n - 1)
    # the searchsorted clamp may land on a dangling node; redirect those few to a non-dangling one
    bad = dang[src]
    if bad.any():
        src[bad] = np.flatnonzero(~dang)[0]
    keep = src != dst
    src, dst = src[keep], dst[keep]
    src, dst = perm[src].astype(np.int32), perm[dst].astype(np.int32)
    if self_loops:
        pos = np.sort(rng.integers(0, len(src), self_loops))
        loops = rng.integers(0, n, self_loops).astype(np.int32)
        src = np.insert(src, pos, loops)
        dst = np.insert(dst, pos, loops)
    return src, dst


def csr_from_edges_np(n, src, dst):
    """numpy restatement of graph.h:152-160 (stable, self loops dropped) used to feed the oracle."""
    keep = src != dst
    s, d = src[keep], dst[keep]
    out_ptr = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(s, minlength=n), out=out_ptr[1:])
    in_ptr = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(d, minlength=n), out=in_ptr[1:])
    out_col = d[np.argsort(s, kind="stable")].astype(np.int32)
    in_col = s[np.argsort(d, kind="stable")].astype(np.int32)
    return out_ptr, out_col, in_ptr, in_col


def write_dataset(folder, n, m_decl, src, dst, queries=None):
    os.makedirs(folder, exist_ok=True)
    with open(os.path.join(folder, "attribute.txt"), "w") as f:
        f.write("n=%d\nm=%d\n" % (n, m_decl))
    np.savetxt(os.path.join(folder, "graph.txt"), np.c_[src, dst], fmt="%d")
    if queries is not None:
        np.savetxt(os.path.join(folder, "ssquery.txt"), np.asarray(queries), fmt="%d")


class Graph:
    def __init__(self, n, src, dst, m_decl=None):
        self.n = int(n)
        self.src = np.ascontiguousarray(src, np.int32)
        self.dst = np.ascontiguousarray(dst, np.int32)
        self.out_ptr, self.out_col, self.in_ptr, self.in_col = csr_from_edges_np(n, self.src, self.dst)
        self.m_decl = int(len(self.src) if m_decl is None else m_decl)
        self.deg = np.diff(self.out_ptr).astype(np.int64)

    @classmethod
    def synth(cls, n, m, seed=42, **kw):
        s, d = synth_edges(n, m, seed, **kw)
        return cls(n, s, d)


# ----------------------------------------------------------------------------- oracle binding
class Oracle:
    def __init__(self, g: Graph, seed=1):
        self.lib = L = C.CDLL(build_oracle())
        self.g = g
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_longlong, c_lp, c_ip, c_lp, c_ip]
        for name in ("orc_forward_push_fifo", "orc_push_topk_round", "orc_forward_push_sync", "orc_fora_query_basic",
                     "orc_kth_ppr", "orc_topk_ppr"):
            getattr(L, name).restype = C.c_double
        L.orc_forward_push_fifo.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.orc_push_topk_begin.argtypes = [C.c_void_p, C.c_int]
        L.orc_push_topk_round.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.orc_push_topk_candidates.argtypes = [C.c_void_p, c_ip]
        L.orc_forward_push_sync.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int]
        L.orc_reverse_push.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int]
        L.orc_seed.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_set_params.argtypes = [C.c_void_p] + [C.c_double] * 7 + [C.c_int] * 3 + [C.c_uint]
        L.orc_init_state.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.orc_get_fwd.argtypes = [C.c_void_p, c_dp, c_dp]
        L.orc_get_bwd.argtypes = [C.c_void_p, c_dp, c_dp]
        L.orc_get_ppr.argtypes = [C.c_void_p, c_dp]
        L.orc_set_fwd.argtypes = [C.c_void_p, c_dp, c_dp]
        L.orc_get_residue_occur.argtypes = [C.c_void_p, c_ip]
        L.orc_get_reserve_occur.argtypes = [C.c_void_p, c_ip]
        L.orc_get_counters.argtypes = [C.c_void_p, c_up]
        L.orc_reset_counters.argtypes = [C.c_void_p]
        L.orc_random_walks.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_int, c_ip]
        L.orc_compute_ppr_with_reserve.argtypes = [C.c_void_p]
        for name in ("orc_compute_ppr_with_fwdidx", "orc_compute_ppr_with_fwdidx_opt", "orc_compute_ppr_with_fwdidx_topk",
                     "orc_compute_ppr_with_fwdidx_topk_with_bound"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_double]
        L.orc_walk_plan.restype = C.c_longlong
        L.orc_walk_plan.argtypes = [C.c_void_p, C.c_double, C.c_int, c_ip, c_up, c_dp]
        L.orc_fora_query_basic.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_double] * 4 + [c_dp]
        L.orc_montecarlo_query.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_bippr_query.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_fwdpush_query.argtypes = [C.c_void_p, C.c_int]
        L.orc_fora_query_topk_new.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_fora_query_topk_with_bound.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_kth_ppr.argtypes = [C.c_void_p, C.c_uint]
        L.orc_topk_ppr.argtypes = [C.c_void_p, C.c_uint, c_ip, c_dp]
        L.orc_index_info.restype = C.c_ulonglong
        L.orc_index_info.argtypes = [C.c_void_p, c_up, c_up]
        L.orc_index_build.argtypes = [C.c_void_p, c_up, c_up, c_ip]
        L.orc_index_set.argtypes = [C.c_void_p, c_up, c_up, c_ip]
        L.orc_power_iteration.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp]
        L.orc_destroy.argtypes = [C.c_void_p]
        self.h = L.orc_create(g.n, g.m_decl, _p(g.out_ptr, c_lp), _p(g.out_col, c_ip), _p(g.in_ptr, c_lp), _p(g.in_col, c_ip))
        L.orc_seed(self.h, seed)
        self.alpha = 0.2
        self.delta = self.pfail = 1.0 / g.n
        self._keep = []

    def __del__(self):
        try:
            self.lib.orc_destroy(self.h)
        except Exception:
            pass

    # --- parameters
    def setting(self, which, epsilon, opt=0, rmax_scale=1.0, delta=None, pfail=None):
        L, g = self.lib, self.g
        delta = self.delta if delta is None else delta
        pfail = self.pfail if pfail is None else pfail
        rmax, omega = C.c_double(0), C.c_double(0)
        D = C.c_double
        if which == "fora":
            L.orc_fora_setting(C.c_longlong(g.m_decl), D(epsilon), D(delta), D(pfail), D(self.alpha), opt, D(rmax_scale), C.byref(rmax), C.byref(omega))
        elif which == "fora_topk":
            L.orc_fora_topk_setting(C.c_longlong(g.m_decl), D(epsilon), D(delta), D(pfail), D(rmax_scale), C.byref(rmax), C.byref(omega))
        elif which == "montecarlo":
            L.orc_montecarlo_setting(D(epsilon), D(delta), D(pfail), C.byref(omega))
        elif which == "bippr":
            L.orc_bippr_setting(C.c_longlong(g.m_decl), D(epsilon), D(delta), D(pfail), D(rmax_scale), C.byref(rmax), C.byref(omega))
        elif which == "fwdpush":
            L.orc_fwdpush_setting(g.n, C.c_longlong(g.m_decl), D(epsilon), D(delta), D(rmax_scale), C.byref(rmax))
        return rmax.value, omega.value

    def set_params(self, epsilon, rmax, omega, opt=0, balanced=0, with_idx=0, k=500, rmax_scale=1.0, delta=None, pfail=None):
        delta = self.delta if delta is None else delta
        pfail = self.pfail if pfail is None else pfail
        self.lib.orc_set_params(self.h, self.alpha, epsilon, delta, pfail, rmax, omega, rmax_scale, opt, balanced, with_idx, k)

    def init_state(self, nil=-1.0, topk_mode=0):
        self.lib.orc_init_state(self.h, nil, topk_mode)

    def seed(self, s):
        self.lib.orc_seed(self.h, s)

    # --- state
    def fwd(self):
        r, q = np.zeros(self.g.n), np.zeros(self.g.n)
        self.lib.orc_get_fwd(self.h, _p(r, c_dp), _p(q, c_dp))
        return r, q

    def bwd(self):
        r, q = np.zeros(self.g.n), np.zeros(self.g.n)
        self.lib.orc_get_bwd(self.h, _p(r, c_dp), _p(q, c_dp))
        return r, q

    def set_fwd(self, reserve, residue):
        reserve = np.ascontiguousarray(reserve, np.float64)
        residue = np.ascontiguousarray(residue, np.float64)
        self.lib.orc_set_fwd(self.h, _p(reserve, c_dp), _p(residue, c_dp))

    def ppr(self):
        p = np.zeros(self.g.n)
        self.lib.orc_get_ppr(self.h, _p(p, c_dp))
        return p

    def residue_occur(self):
        k = np.zeros(4 * self.g.n + 16, np.int32)
        c = self.lib.orc_get_residue_occur(self.h, _p(k, c_ip))
        return k[:c].copy()

    def reserve_occur(self):
        k = np.zeros(4 * self.g.n + 16, np.int32)
        c = self.lib.orc_get_reserve_occur(self.h, _p(k, c_ip))
        return k[:c].copy()

    def counters(self):
        o = np.zeros(8, np.uint64)
        self.lib.orc_get_counters(self.h, _p(o, c_up))
        return dict(zip(("total_rw", "hit_idx", "walk_hops", "edges_pushed", "vertices_pushed", "push_levels", "topk_iters", "rounds"), map(int, o)))

    def reset_counters(self):
        self.lib.orc_reset_counters(self.h)

    # --- algorithms
    def push_fifo(self, s, rmax, init=1.0):
        return self.lib.orc_forward_push_fifo(self.h, s, rmax, init)

    def push_topk_begin(self, s):
        self.lib.orc_push_topk_begin(self.h, s)

    def push_topk_round(self, s, rmax, lowest):
        return self.lib.orc_push_topk_round(self.h, s, rmax, lowest)

    def push_topk_candidates(self):
        k = np.zeros(self.g.n + 16, np.int32)
        c = self.lib.orc_push_topk_candidates(self.h, _p(k, c_ip))
        return k[:c].copy()

    def push_sync(self, s, rmax, fresh=1, seed_all=0):
        return self.lib.orc_forward_push_sync(self.h, s, rmax, fresh, seed_all)

    def reverse_push(self, t, rmax, init=1.0, sync=0):
        self.lib.orc_reverse_push(self.h, t, rmax, init, sync)

    def walks(self, start, count, no_zero_hop=0):
        d = np.zeros(count, np.int32)
        self.lib.orc_random_walks(self.h, start, count, no_zero_hop, _p(d, c_ip))
        return d

    def walk_plan(self, rsum, opt):
        n = 4 * self.g.n + 16
        keys, cnt, inc = np.zeros(n, np.int32), np.zeros(n, np.uint64), np.zeros(n)
        c = self.lib.orc_walk_plan(self.h, rsum, opt, _p(keys, c_ip), _p(cnt, c_up), _p(inc, c_dp))
        return keys[:c].copy(), cnt[:c].copy(), inc[:c].copy()

    def compute_ppr(self, which, rsum=0.0):
        fn = {"reserve": None, "fwdidx": "orc_compute_ppr_with_fwdidx", "opt": "orc_compute_ppr_with_fwdidx_opt",
              "topk": "orc_compute_ppr_with_fwdidx_topk", "bound": "orc_compute_ppr_with_fwdidx_topk_with_bound"}[which]
        if fn is None:
            self.lib.orc_compute_ppr_with_reserve(self.h)
        else:
            getattr(self.lib, fn)(self.h, rsum)

    def fora_query(self, s, balanced_mode=0, sync_push=0, walk_cost=4e-7, c_edge=0.0, c_vertex=0.0, c_level=0.0):
        fr = C.c_double(0)
        rsum = self.lib.orc_fora_query_basic(self.h, s, balanced_mode, sync_push, walk_cost, c_edge, c_vertex, c_level, C.byref(fr))
        return rsum, fr.value

    def montecarlo_query(self, s, topk_variant=0):
        self.lib.orc_montecarlo_query(self.h, s, topk_variant)

    def bippr_query(self, s, topk_variant=0, sync_push=0):
        self.lib.orc_bippr_query(self.h, s, topk_variant, sync_push)

    def fwdpush_query(self, s):
        self.lib.orc_fwdpush_query(self.h, s)

    def fora_topk_new(self, s, sync_push=0):
        self.lib.orc_fora_query_topk_new(self.h, s, sync_push)

    def fora_topk_with_bound(self, s, sync_push=0):
        self.lib.orc_fora_query_topk_with_bound(self.h, s, sync_push)

    def kth_ppr(self, k):
        return self.lib.orc_kth_ppr(self.h, k)

    def topk_ppr(self, k):
        nodes, vals = np.zeros(k, np.int32), np.zeros(k)
        self.lib.orc_topk_ppr(self.h, k, _p(nodes, c_ip), _p(vals, c_dp))
        return nodes, vals

    def precision(self, k, est_nodes, est_vals, ex_nodes, ex_vals):
        est_nodes = np.ascontiguousarray(est_nodes, np.int32); est_vals = np.ascontiguousarray(est_vals, np.float64)
        ex_nodes = np.ascontiguousarray(ex_nodes, np.int32); ex_vals = np.ascontiguousarray(ex_vals, np.float64)
        p, r = C.c_double(0), C.c_double(0)
        self.lib.orc_precision.argtypes = [C.c_uint, C.c_int, c_ip, c_dp, C.c_int, c_ip, c_dp, c_dp, c_dp]
        self.lib.orc_precision(k, len(est_nodes), _p(est_nodes, c_ip), _p(est_vals, c_dp), len(ex_nodes), _p(ex_nodes, c_ip),
                               _p(ex_vals, c_dp), C.byref(p), C.byref(r))
        return p.value, r.value

    def index_info(self):
        off, cnt = np.zeros(self.g.n, np.uint64), np.zeros(self.g.n, np.uint64)
        total = self.lib.orc_index_info(self.h, _p(off, c_up), _p(cnt, c_up))
        return off, cnt, int(total)

    def index_build(self, off, cnt):
        dest = np.zeros(int(cnt.sum()), np.int32)
        self.lib.orc_index_build(self.h, _p(off, c_up), _p(cnt, c_up), _p(dest, c_ip))
        return dest

    def index_set(self, off, cnt, dest):
        self._keep = [np.ascontiguousarray(off, np.uint64), np.ascontiguousarray(cnt, np.uint64), np.ascontiguousarray(dest, np.int32)]
        self.lib.orc_index_set(self.h, _p(self._keep[0], c_up), _p(self._keep[1], c_up), _p(self._keep[2], c_ip))

    def power_iteration(self, s, iters=100):
        p = np.zeros(self.g.n)
        self.lib.orc_power_iteration(self.h, s, iters, _p(p, c_dp))
        return p


# ----------------------------------------------------------------------------- reference binding
class Reference:
    """One fresh copy of libfora_ref.so per instance: the reference keeps process-wide globals and
    function-local statics, so every (graph, config) gets its own loaded image."""
    _in_place_used = False

    def __init__(self, g: Graph = None, folder=None, epsilon=0.5, opt=0, balanced=0, with_idx=0, rmax_scale=1.0, k=500):
        assert have_reference(), "oracle/_ref/libfora_ref.so not built (make -C oracle ref)"
        # the first instance of a process maps oracle/_ref/libfora_ref.so where it lies (so a loader trace of the process shows the
        # reference itself); later instances need their own image of the reference's globals and map a private copy
        if not Reference._in_place_used:
            Reference._in_place_used = True
            so = REF_SO
        else:
            self._tmp = tempfile.mkdtemp(prefix="fora_ref_so_")
            so = os.path.join(self._tmp, "libfora_ref.so")
            shutil.copy(REF_SO, so)
        self.lib = L = C.CDLL(so)
        L.ref_graph_from_csr.argtypes = [C.c_int, C.c_longlong, c_lp, c_ip, c_lp, c_ip]
        L.ref_graph_load_dir.argtypes = [C.c_char_p]
        L.ref_graph_m.restype = C.c_longlong
        L.ref_graph_num_out_edges.restype = C.c_longlong
        L.ref_graph_dump.argtypes = [c_lp, c_ip, c_lp, c_ip]
        L.ref_config.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint]
        L.ref_set_delta_pfail.argtypes = [C.c_double, C.c_double]
        L.ref_setting.argtypes = [C.c_int, c_dp, c_dp]
        L.ref_get_params.argtypes = [c_dp]
        L.ref_set_rmax_omega.argtypes = [C.c_double, C.c_double]
        L.ref_dump_fwd.argtypes = [c_dp, c_ip, c_ip, c_dp, c_ip, c_ip]
        L.ref_dump_bwd.argtypes = [c_dp, c_ip, c_dp, c_ip]
        L.ref_dump_ppr.argtypes = [c_dp]
        L.ref_counters.argtypes = [c_up, c_up]
        L.ref_timer_used.restype = C.c_double
        L.ref_timer_used.argtypes = [C.c_int]
        L.ref_forward_push.restype = C.c_double
        L.ref_forward_push.argtypes = [C.c_int, C.c_double, C.c_double]
        L.ref_push_topk_begin.argtypes = [C.c_int]
        L.ref_push_topk_round.restype = C.c_double
        L.ref_push_topk_round.argtypes = [C.c_int, C.c_double, C.c_double]
        L.ref_push_topk_candidates.argtypes = [c_ip]
        L.ref_reverse_push.argtypes = [C.c_int, C.c_double]
        L.ref_random_walks.argtypes = [C.c_int, C.c_longlong, C.c_int, c_ip]
        L.ref_compute_ppr.argtypes = [C.c_int, C.c_double]
        L.ref_query.argtypes = [C.c_int, C.c_int]
        L.ref_topk.argtypes = [C.c_int, C.c_int, c_ip, c_dp]
        L.ref_num_iter_topk.restype = C.c_long
        L.ref_set_exact_topk.argtypes = [C.c_int, C.c_int, c_ip, c_dp]
        L.ref_set_topk_pprs.argtypes = [C.c_int, c_ip, c_dp]
        L.ref_compute_precision.argtypes = [C.c_int, c_dp, c_dp]
        L.ref_power_iteration.argtypes = [C.c_int, c_dp]
        L.ref_build_index.restype = C.c_longlong
        L.ref_build_index.argtypes = [C.c_char_p]
        L.ref_load_index.argtypes = [C.c_char_p]
        L.ref_index_size.restype = C.c_longlong
        L.ref_index_dump.argtypes = [c_up, c_up, c_ip]
        L.ref_index_set.argtypes = [C.c_int, c_up, c_up, C.c_longlong, c_ip]
        L.ref_index_file_names.restype = C.c_char_p
        L.ref_index_file_names.argtypes = [C.c_int]
        L.ref_save_exact_topk.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_load_exact_topk.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_get_exact_topk.argtypes = [C.c_int, C.c_int, c_ip, c_dp]
        L.ref_config(epsilon, opt, balanced, with_idx, rmax_scale, k)
        if folder is not None:
            self.n = L.ref_graph_load_dir(folder.encode())
        else:
            self.g = g
            self.n = L.ref_graph_from_csr(g.n, g.m_decl, _p(g.out_ptr, c_lp), _p(g.out_col, c_ip), _p(g.in_ptr, c_lp), _p(g.in_col, c_ip))
        assert self.n > 0

    def __del__(self):
        if getattr(self, "_tmp", None):
            shutil.rmtree(self._tmp, ignore_errors=True)

    def graph_dump(self):
        ne = self.lib.ref_graph_num_out_edges()
        op, ip_ = np.zeros(self.n + 1, np.int64), np.zeros(self.n + 1, np.int64)
        oc, ic = np.zeros(ne, np.int32), np.zeros(ne, np.int32)
        self.lib.ref_graph_dump(_p(op, c_lp), _p(oc, c_ip), _p(ip_, c_lp), _p(ic, c_ip))
        return op, oc, ip_, ic

    def setting(self, which):
        idx = {"fora": 0, "fora_topk": 1, "montecarlo": 2, "bippr": 3, "fwdpush": 4}[which]
        r, o = C.c_double(0), C.c_double(0)
        self.lib.ref_setting(idx, C.byref(r), C.byref(o))
        return r.value, o.value

    def params(self):
        o = np.zeros(6)
        self.lib.ref_get_params(_p(o, c_dp))
        return dict(zip(("alpha", "epsilon", "delta", "pfail", "rmax", "omega"), o))

    def init_query_state(self):
        self.lib.ref_init_query_state()

    def init_topk_state(self, algo="fora"):
        if algo == "fora":
            self.lib.ref_init_topk_state_fora()
        else:
            self.lib.ref_init_topk_state_other({"montecarlo": 2, "bippr": 3, "fwdpush": 4}[algo])

    def fwd(self):
        n = self.n
        r, q = np.zeros(n), np.zeros(n)
        ro, qo = np.zeros(4 * n + 16, np.int32), np.zeros(4 * n + 16, np.int32)
        nr, nq = C.c_int(0), C.c_int(0)
        self.lib.ref_dump_fwd(_p(r, c_dp), _p(ro, c_ip), C.byref(nr), _p(q, c_dp), _p(qo, c_ip), C.byref(nq))
        return r, q, ro[: nr.value].copy(), qo[: nq.value].copy()

    def bwd(self):
        r, q = np.zeros(self.n), np.zeros(self.n)
        a, b = C.c_int(0), C.c_int(0)
        self.lib.ref_dump_bwd(_p(r, c_dp), C.byref(a), _p(q, c_dp), C.byref(b))
        return r, q

    def ppr(self):
        p = np.zeros(self.n)
        self.lib.ref_dump_ppr(_p(p, c_dp))
        return p

    def counters(self):
        a, b = C.c_ulonglong(0), C.c_ulonglong(0)
        self.lib.ref_counters(C.byref(a), C.byref(b))
        return a.value, b.value

    def push(self, s, rmax, init=1.0):
        return self.lib.ref_forward_push(s, rmax, init)

    def push_topk_begin(self, s):
        self.lib.ref_push_topk_begin(s)

    def push_topk_round(self, s, rmax, lowest):
        return self.lib.ref_push_topk_round(s, rmax, lowest)

    def push_topk_candidates(self):
        k = np.zeros(self.n + 16, np.int32)
        c = self.lib.ref_push_topk_candidates(_p(k, c_ip))
        return k[:c].copy()

    def reverse_push(self, t, init=1.0):
        self.lib.ref_reverse_push(t, init)

    def walks(self, start, count, no_zero_hop=0):
        d = np.zeros(count, np.int32)
        self.lib.ref_random_walks(start, count, no_zero_hop, _p(d, c_ip))
        return d

    def compute_ppr(self, which, rsum=0.0):
        self.lib.ref_compute_ppr({"fwdidx": 0, "opt": 1, "topk": 2, "bound": 3, "reserve": 4}[which], rsum)

    def query(self, algo, s):
        self.lib.ref_query({"fora": 0, "montecarlo": 2, "bippr": 3, "fwdpush": 4}[algo], s)

    def topk(self, algo, s, k):
        nodes, vals = np.zeros(k, np.int32), np.zeros(k)
        self.lib.ref_topk({"fora": 0, "montecarlo": 2, "bippr": 3, "fwdpush": 4}[algo], s, _p(nodes, c_ip), _p(vals, c_dp))
        return nodes, vals

    def precision(self, v, est_nodes, est_vals, ex_nodes, ex_vals):
        est_nodes = np.ascontiguousarray(est_nodes, np.int32); est_vals = np.ascontiguousarray(est_vals, np.float64)
        ex_nodes = np.ascontiguousarray(ex_nodes, np.int32); ex_vals = np.ascontiguousarray(ex_vals, np.float64)
        self.lib.ref_set_exact_topk(v, len(ex_nodes), _p(ex_nodes, c_ip), _p(ex_vals, c_dp))
        self.lib.ref_set_topk_pprs(len(est_nodes), _p(est_nodes, c_ip), _p(est_vals, c_dp))
        p, r = C.c_double(0), C.c_double(0)
        self.lib.ref_compute_precision(v, C.byref(p), C.byref(r))
        return p.value, r.value

    def power_iteration(self, s):
        p = np.zeros(self.n)
        self.lib.ref_power_iteration(s, _p(p, c_dp))
        return p

    def build_index(self, folder):
        if not folder.endswith("/"):
            folder += "/"
        total = self.lib.ref_build_index(folder.encode())
        off, cnt = np.zeros(self.n, np.uint64), np.zeros(self.n, np.uint64)
        dest = np.zeros(total, np.int32)
        self.lib.ref_index_dump(_p(off, c_up), _p(cnt, c_up), _p(dest, c_ip))
        return off, cnt, dest

    def index_set(self, off, cnt, dest):
        off = np.ascontiguousarray(off, np.uint64); cnt = np.ascontiguousarray(cnt, np.uint64)
        dest = np.ascontiguousarray(dest, np.int32)
        self.lib.ref_index_set(self.n, _p(off, c_up), _p(cnt, c_up), len(dest), _p(dest, c_ip))

    def timer(self, i):
        return self.lib.ref_timer_used(i)
