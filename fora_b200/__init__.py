"""fora_b200 -- B200-native FORA engine (hand-written sm_100a CUDA behind a C ABI).

This module is the thin ctypes binding of ``include/fora_b200.h`` used by the tests and the
benchmark; the reference-facing host program is the C++ ``./fora`` CLI under ``fora_b200/host``.
The library has no CPU fallback: loading fails loudly when ``libfora_b200.so`` has not been
built (``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C fora_b200``) and
``Engine()`` raises when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfora_b200.so")

ALGO = {"fora": 0, "fwdpush": 1, "montecarlo": 2, "bippr": 3}
SETTING = {"fora": 0, "fora_topk": 1, "montecarlo": 2, "bippr": 3, "fwdpush": 4}

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_lp = C.POINTER(C.c_int64)
c_up = C.POINTER(C.c_uint64)


class ForaError(RuntimeError):
    pass


class Params(C.Structure):
    """mirror of struct fora_params (== the reference's global `config`, config.h:86-138)"""
    _fields_ = [("alpha", C.c_double), ("epsilon", C.c_double), ("delta", C.c_double), ("pfail", C.c_double),
                ("rmax", C.c_double), ("omega", C.c_double), ("rmax_scale", C.c_double),
                ("opt", C.c_int32), ("balanced", C.c_int32), ("with_idx", C.c_int32), ("k", C.c_uint32),
                ("cost_walk", C.c_double), ("cost_edge", C.c_double), ("cost_vertex", C.c_double), ("cost_level", C.c_double)]


class QueryStat(C.Structure):
    _fields_ = [("rsum", C.c_double), ("final_rmax", C.c_double), ("n_walks", C.c_uint64), ("n_idx_hits", C.c_uint64),
                ("walk_hops", C.c_uint64), ("edges_pushed", C.c_uint64), ("vertices_pushed", C.c_uint64),
                ("push_levels", C.c_uint64), ("push_rounds", C.c_uint64), ("n_sources", C.c_uint64)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class SplitTiming(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("push_ms", C.c_float), ("bcast_ms", C.c_float), ("walk_ms", C.c_float), ("reduce_ms", C.c_float),
                ("n_gpus", C.c_uint32), ("bcast_bytes", C.c_uint64), ("reduce_bytes", C.c_uint64)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class BatchTiming(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("push_ms", C.c_float), ("walk_ms", C.c_float), ("plan_ms", C.c_float),
                ("topk_ms", C.c_float), ("copy_ms", C.c_float), ("kernel_launches", C.c_uint64),
                ("push_kernel_ms", C.c_float), ("walk_kernel_ms", C.c_float),
                ("push_kernel_launches", C.c_uint64), ("walk_kernel_launches", C.c_uint64)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


_lib = None
_host = None
HOST_LIB_PATH = os.path.join(_HERE, "libfora_host.so")


def _bind_host(L):
    L.fora_host_read_attribute.argtypes = [C.c_char_p, c_ip, c_lp]
    L.fora_host_read_edges.restype = C.c_int64
    L.fora_host_read_edges.argtypes = [C.c_char_p, C.c_int32, c_ip, c_ip]
    L.fora_host_csr_from_edges.argtypes = [C.c_int32, C.c_int64, c_ip, c_ip, c_lp, c_ip, c_lp, c_ip]
    L.fora_host_synth_edges.restype = C.c_int64
    L.fora_host_synth_edges.argtypes = [C.c_int32, C.c_int64, C.c_uint64, C.c_double, C.c_double, c_ip, c_ip]
    L.fora_host_setting.argtypes = [C.c_int, C.c_int32, C.c_int64] + [C.c_double] * 4 + [C.c_int, C.c_double, c_dp, c_dp]
    return L


def host_lib():
    """The host-only helpers of the ABI (loader, CSR from edges, synthetic graphs, *_setting): libfora_host.so, the same
    host_util.cpp without the GPU engine -- what a process that must not map the engine uses (bench.py --impl reference).
    A process that has already loaded libfora_b200.so uses that copy."""
    global _host
    if _host is None:
        if _lib is not None:
            _host = _lib
        elif os.path.exists(HOST_LIB_PATH):
            _host = _bind_host(C.CDLL(HOST_LIB_PATH))
        else:
            _host = lib()
    return _host


def lib():
    """Load libfora_b200.so (in-tree).  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ForaError("%s is missing: build it with `make -C fora_b200` (there is no CPU fallback)" % LIB_PATH)
    L = _bind_host(C.CDLL(LIB_PATH))
    vp = C.c_void_p
    L.fora_ctx_create.argtypes = [C.c_int, C.c_uint64, C.POINTER(vp)]
    L.fora_ctx_destroy.argtypes = [vp]
    L.fora_last_error.restype = C.c_char_p
    L.fora_last_error.argtypes = [vp]
    L.fora_ctx_set_stream.argtypes = [vp, vp]
    L.fora_ctx_set_slots.argtypes = [vp, C.c_int]
    L.fora_ctx_sync.argtypes = [vp]
    L.fora_ctx_set_query_base.argtypes = [vp, C.c_uint64]
    L.fora_ctx_set_shared_walks.argtypes = [vp, C.c_int]
    L.fora_graph_upload.argtypes = [vp, C.c_int32, C.c_int64, c_lp, c_ip, c_lp, c_ip]
    L.fora_graph_build_from_edges.argtypes = [vp, C.c_int32, C.c_int64, c_ip, c_ip, C.c_int64, C.c_int]
    L.fora_graph_download_csr.argtypes = [vp, c_lp, c_ip, c_lp, c_ip]
    L.fora_graph_num_edges.restype = C.c_int64
    L.fora_graph_num_edges.argtypes = [vp]
    L.fora_params_set.argtypes = [vp, C.POINTER(Params)]
    L.fora_params_get.argtypes = [vp, C.POINTER(Params)]
    L.fora_push_only.argtypes = [vp, C.c_int32, C.c_double, c_dp, c_dp, c_dp, C.POINTER(QueryStat)]
    L.fora_push_begin.argtypes = [vp, C.c_int32]
    L.fora_push_round.argtypes = [vp, C.c_double, c_dp, c_dp, c_dp, C.POINTER(QueryStat)]
    L.fora_reverse_push.argtypes = [vp, C.c_int32, C.c_double, c_dp, c_dp]
    L.fora_random_walks.argtypes = [vp, C.c_int32, C.c_int64, C.c_int, c_ip, c_up]
    L.fora_compute_ppr.argtypes = [vp, c_dp, c_dp, C.c_double, c_dp, C.POINTER(QueryStat)]
    L.fora_query_batch.argtypes = [vp, C.c_int, c_ip, C.c_int32, c_dp, C.POINTER(QueryStat), C.POINTER(BatchTiming)]
    L.fora_query_batch_sparse.argtypes = [vp, C.c_int, c_ip, C.c_int32, C.c_double, C.c_uint64, C.c_uint64, c_ip, c_dp, c_up, C.POINTER(QueryStat), C.POINTER(BatchTiming)]
    L.fora_query_batch_device.argtypes = [vp, C.c_int, vp, C.c_int32, C.POINTER(QueryStat), C.POINTER(BatchTiming)]
    L.fora_device_ppr.restype = vp
    L.fora_device_ppr.argtypes = [vp, C.c_int]
    L.fora_prepare_slots.argtypes = [vp]
    L.fora_device_reserve.restype = vp
    L.fora_device_reserve.argtypes = [vp, C.c_int]
    L.fora_device_residue.restype = vp
    L.fora_device_residue.argtypes = [vp, C.c_int]
    L.fora_device_to_original.argtypes = [vp, vp, vp]
    L.fora_compute_ppr_part_device.argtypes = [vp, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(QueryStat)]
    L.fora_topk_batch.argtypes = [vp, C.c_int, c_ip, C.c_int32, C.c_uint32, c_ip, c_dp, c_ip, C.POINTER(QueryStat), C.POINTER(BatchTiming)]
    L.fora_topk_of.argtypes = [vp, c_dp, C.c_uint32, c_ip, c_dp]
    L.fora_index_info.argtypes = [vp, c_up, c_up, c_up]
    L.fora_index_build.argtypes = [vp, c_up, c_up, C.c_int32, C.c_int32, c_ip]
    L.fora_index_upload.argtypes = [vp, c_up, c_up, c_ip, C.c_uint64]
    L.fora_index_build_stat.argtypes = [vp, c_up, c_up, c_dp]
    L.fora_power_iteration.argtypes = [vp, C.c_int32, C.c_int, c_dp]
    L.fora_group_create.argtypes = [C.c_int, c_ip, C.c_uint64, C.POINTER(vp)]
    L.fora_group_destroy.argtypes = [vp]
    L.fora_group_size.argtypes = [vp]
    L.fora_group_ctx.restype = vp
    L.fora_group_ctx.argtypes = [vp, C.c_int]
    L.fora_group_last_error.restype = C.c_char_p
    L.fora_group_last_error.argtypes = [vp]
    L.fora_group_query_split.argtypes = [vp, C.c_int32, C.c_uint32, c_dp, C.POINTER(QueryStat), C.POINTER(SplitTiming)]
    _lib = L
    return L


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


# ------------------------------------------------------------------------------- host helpers
def read_attribute(path):
    n, m = C.c_int32(0), C.c_int64(0)
    rc = host_lib().fora_host_read_attribute(path.encode(), C.byref(n), C.byref(m))
    if rc:
        raise ForaError("cannot read %s (rc=%d)" % (path, rc))
    return n.value, m.value


def read_edges(path, n):
    L = host_lib()
    cnt = L.fora_host_read_edges(path.encode(), n, None, None)
    if cnt < 0:
        raise ForaError("cannot read %s (rc=%d)" % (path, cnt))
    src, dst = np.empty(cnt, np.int32), np.empty(cnt, np.int32)
    L.fora_host_read_edges(path.encode(), n, _p(src, c_ip), _p(dst, c_ip))
    return src, dst


def csr_from_edges(n, src, dst, with_in=True):
    src = np.ascontiguousarray(src, np.int32)
    dst = np.ascontiguousarray(dst, np.int32)
    kept = int((src != dst).sum())
    out_ptr, out_col = np.empty(n + 1, np.int64), np.empty(max(kept, 1), np.int32)
    in_ptr = np.empty(n + 1, np.int64) if with_in else None
    in_col = np.empty(max(kept, 1), np.int32) if with_in else None
    rc = host_lib().fora_host_csr_from_edges(n, len(src), _p(src, c_ip), _p(dst, c_ip), _p(out_ptr, c_lp), _p(out_col, c_ip),
                                        _p(in_ptr, c_lp), _p(in_col, c_ip))
    if rc:
        raise ForaError("csr_from_edges rc=%d" % rc)
    if with_in:
        return out_ptr, out_col[:kept], in_ptr, in_col[:kept]
    return out_ptr, out_col[:kept], None, None


def synth_edges(n, m, seed=42, exponent=2.3, dangling_frac=0.03):
    src, dst = np.empty(m, np.int32), np.empty(m, np.int32)
    rc = host_lib().fora_host_synth_edges(n, m, seed, exponent, dangling_frac, _p(src, c_ip), _p(dst, c_ip))
    if rc < 0:
        raise ForaError("synth_edges rc=%d" % rc)
    return src, dst


def setting(which, n, m, epsilon, delta=None, pfail=None, alpha=0.2, opt=0, rmax_scale=1.0):
    """*_setting of algo.h:442-496; delta/pfail default to init_parameter's 1/n (graph.h:177-178)."""
    delta = 1.0 / n if delta is None else delta
    pfail = 1.0 / n if pfail is None else pfail
    rmax, omega = C.c_double(0), C.c_double(0)
    rc = host_lib().fora_host_setting(SETTING[which], n, m, epsilon, delta, pfail, alpha, opt, rmax_scale, C.byref(rmax), C.byref(omega))
    if rc:
        raise ForaError("setting rc=%d" % rc)
    return rmax.value, omega.value


# ------------------------------------------------------------------------------- engine
class Engine:
    """One context on one GPU (fora_ctx)."""

    def __init__(self, device=0, seed=1, slots=None):
        self.L = lib()
        self.h = C.c_void_p()
        rc = self.L.fora_ctx_create(device, seed, C.byref(self.h))
        if rc:
            raise ForaError("fora_ctx_create failed: %s" % self.L.fora_last_error(None).decode())
        self.n = 0
        self.m_decl = 0
        self.params = Params()
        if slots:
            self.set_slots(slots)

    def close(self):
        if getattr(self, "h", None):
            self.L.fora_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise ForaError("rc=%d: %s" % (rc, self.L.fora_last_error(self.h).decode()))

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.L.fora_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_slots(self, slots):
        self._ck(self.L.fora_ctx_set_slots(self.h, slots))

    def sync(self):
        self._ck(self.L.fora_ctx_sync(self.h))

    def set_query_base(self, first_query_index):
        """global list index of the first query of the next batch call (Philox key), see include/fora_b200.h"""
        self._ck(self.L.fora_ctx_set_query_base(self.h, int(first_query_index)))

    def set_shared_walks(self, on=True):
        """opt-in: the queries of one wave draw their walks from one pool (per-wave virtual walk index), see include/fora_b200.h"""
        self._ck(self.L.fora_ctx_set_shared_walks(self.h, 1 if on else 0))

    # --- graph
    def upload_graph(self, n, m_decl, out_ptr, out_col, in_ptr=None, in_col=None):
        out_ptr = np.ascontiguousarray(out_ptr, np.int64)
        out_col = np.ascontiguousarray(out_col, np.int32)
        if in_ptr is not None:
            in_ptr = np.ascontiguousarray(in_ptr, np.int64)
            in_col = np.ascontiguousarray(in_col, np.int32)
        self._ck(self.L.fora_graph_upload(self.h, n, m_decl, _p(out_ptr, c_lp), _p(out_col, c_ip), _p(in_ptr, c_lp), _p(in_col, c_ip)))
        self.n, self.m_decl = n, m_decl

    def build_graph_from_edges(self, n, m_decl, src, dst, with_in=True):
        """K0: CSR built on the device from the edge list in file order"""
        src = np.ascontiguousarray(src, np.int32)
        dst = np.ascontiguousarray(dst, np.int32)
        self._ck(self.L.fora_graph_build_from_edges(self.h, n, m_decl, _p(src, c_ip), _p(dst, c_ip), len(src), int(with_in)))
        self.n, self.m_decl = n, m_decl

    def download_csr(self, with_in=True):
        ne = self.L.fora_graph_num_edges(self.h)
        op, oc = np.empty(self.n + 1, np.int64), np.empty(ne, np.int32)
        ip_ = np.empty(self.n + 1, np.int64) if with_in else None
        ic = np.empty(ne, np.int32) if with_in else None
        self._ck(self.L.fora_graph_download_csr(self.h, _p(op, c_lp), _p(oc, c_ip), _p(ip_, c_lp), _p(ic, c_ip)))
        return op, oc, ip_, ic

    # --- parameters
    def set_params(self, epsilon, rmax, omega, alpha=0.2, delta=None, pfail=None, rmax_scale=1.0, opt=0, balanced=0,
                   with_idx=0, k=500, cost_walk=0.0, cost_edge=0.0, cost_vertex=0.0, cost_level=0.0):
        p = Params(alpha, epsilon, 1.0 / self.n if delta is None else delta, 1.0 / self.n if pfail is None else pfail,
                   rmax, omega, rmax_scale, opt, balanced, with_idx, k, cost_walk, cost_edge, cost_vertex, cost_level)
        self._ck(self.L.fora_params_set(self.h, C.byref(p)))
        self._ck(self.L.fora_params_get(self.h, C.byref(self.params)))

    def configure(self, algo="fora", epsilon=0.5, opt=0, balanced=0, with_idx=0, rmax_scale=1.0, k=500, **cost):
        """what query() does before its loop: *_setting then the per-algo state (query.h:1429-1511)"""
        which = {"fora": "fora", "fwdpush": "fwdpush", "montecarlo": "montecarlo", "bippr": "bippr"}[algo]
        rmax, omega = setting(which, self.n, self.m_decl, epsilon, opt=opt, rmax_scale=rmax_scale)
        self.set_params(epsilon, rmax, omega, opt=opt, balanced=balanced, with_idx=with_idx, rmax_scale=rmax_scale, k=k, **cost)
        return rmax, omega

    # --- push
    def push_only(self, source, rmax):
        reserve, residue = np.empty(self.n), np.empty(self.n)
        rsum, st = C.c_double(0), QueryStat()
        self._ck(self.L.fora_push_only(self.h, source, rmax, _p(reserve, c_dp), _p(residue, c_dp), C.byref(rsum), C.byref(st)))
        return reserve, residue, rsum.value, st.as_dict()

    def push_begin(self, source):
        self._ck(self.L.fora_push_begin(self.h, source))

    def push_round(self, rmax):
        reserve, residue = np.empty(self.n), np.empty(self.n)
        rsum, st = C.c_double(0), QueryStat()
        self._ck(self.L.fora_push_round(self.h, rmax, _p(reserve, c_dp), _p(residue, c_dp), C.byref(rsum), C.byref(st)))
        return reserve, residue, rsum.value, st.as_dict()

    def reverse_push(self, target, rmax):
        reserve, residue = np.empty(self.n), np.empty(self.n)
        self._ck(self.L.fora_reverse_push(self.h, target, rmax, _p(reserve, c_dp), _p(residue, c_dp)))
        return reserve, residue

    # --- walks
    def random_walks(self, start, count, no_zero_hop=0):
        dest = np.empty(count, np.int32)
        hops = C.c_uint64(0)
        self._ck(self.L.fora_random_walks(self.h, start, count, no_zero_hop, _p(dest, c_ip), C.byref(hops)))
        return dest, hops.value

    def compute_ppr(self, reserve, residue, rsum):
        reserve = np.ascontiguousarray(reserve, np.float64)
        residue = np.ascontiguousarray(residue, np.float64)
        ppr, st = np.empty(self.n), QueryStat()
        self._ck(self.L.fora_compute_ppr(self.h, _p(reserve, c_dp), _p(residue, c_dp), rsum, _p(ppr, c_dp), C.byref(st)))
        return ppr, st.as_dict()

    # --- queries
    def query_batch(self, algo, sources, want_ppr=True, out=None):
        sources = np.ascontiguousarray(sources, np.int32)
        nq = len(sources)
        ppr = None
        if want_ppr:
            ppr = out if out is not None else np.empty((nq, self.n))
        stats = (QueryStat * max(nq, 1))()
        tm = BatchTiming()
        self._ck(self.L.fora_query_batch(self.h, ALGO[algo], _p(sources, c_ip), nq, _p(ppr, c_dp), stats, C.byref(tm)))
        return ppr, [stats[i].as_dict() for i in range(nq)], tm.as_dict()

    def query_batch_sparse(self, algo, sources, threshold, cap_per_query, ids=None, vals=None):
        """compacted result: (ids, vals, offsets) with the entries >= threshold of query i at [offsets[i], offsets[i+1])"""
        sources = np.ascontiguousarray(sources, np.int32)
        nq = len(sources)
        if ids is None:
            ids, vals = np.empty(nq * cap_per_query, np.int32), np.empty(nq * cap_per_query)
        off = np.zeros(nq + 1, np.uint64)
        stats = (QueryStat * max(nq, 1))()
        tm = BatchTiming()
        self._ck(self.L.fora_query_batch_sparse(self.h, ALGO[algo], _p(sources, c_ip), nq, threshold, cap_per_query, len(ids), _p(ids, c_ip), _p(vals, c_dp),
                                                _p(off, c_up), stats, C.byref(tm)))
        return ids, vals, off, [stats[i].as_dict() for i in range(nq)], tm.as_dict()

    def query_batch_device(self, algo, d_sources_ptr, nq):
        stats = (QueryStat * max(nq, 1))()
        tm = BatchTiming()
        self._ck(self.L.fora_query_batch_device(self.h, ALGO[algo], C.c_void_p(d_sources_ptr), nq, stats, C.byref(tm)))
        return [stats[i].as_dict() for i in range(nq)], tm.as_dict()

    def prepare_slots(self):
        self._ck(self.L.fora_prepare_slots(self.h))

    def device_reserve_ptr(self, slot=0):
        return self.L.fora_device_reserve(self.h, slot)

    def device_residue_ptr(self, slot=0):
        return self.L.fora_device_residue(self.h, slot)

    def device_to_original(self, d_internal_ptr, d_original_ptr):
        self._ck(self.L.fora_device_to_original(self.h, C.c_void_p(d_internal_ptr), C.c_void_p(d_original_ptr)))

    def compute_ppr_part_device(self, rsum, qid, part, nparts):
        st = QueryStat()
        self._ck(self.L.fora_compute_ppr_part_device(self.h, rsum, qid, part, nparts, C.byref(st)))
        return st.as_dict()

    def device_ppr_ptr(self, slot):
        return self.L.fora_device_ppr(self.h, slot)

    def topk_batch(self, algo, sources, k):
        sources = np.ascontiguousarray(sources, np.int32)
        nq = len(sources)
        nodes, vals, iters = np.zeros((nq, k), np.int32), np.zeros((nq, k)), np.zeros(nq, np.int32)
        stats = (QueryStat * max(nq, 1))()
        tm = BatchTiming()
        self._ck(self.L.fora_topk_batch(self.h, ALGO[algo], _p(sources, c_ip), nq, k, _p(nodes, c_ip), _p(vals, c_dp), _p(iters, c_ip), stats, C.byref(tm)))
        return nodes, vals, iters, [stats[i].as_dict() for i in range(nq)], tm.as_dict()

    def topk_of(self, ppr, k):
        ppr = np.ascontiguousarray(ppr, np.float64)
        nodes, vals = np.zeros(k, np.int32), np.zeros(k)
        self._ck(self.L.fora_topk_of(self.h, _p(ppr, c_dp), k, _p(nodes, c_ip), _p(vals, c_dp)))
        return nodes, vals

    # --- index
    def index_info(self):
        off, cnt = np.empty(self.n, np.uint64), np.empty(self.n, np.uint64)
        total = C.c_uint64(0)
        self._ck(self.L.fora_index_info(self.h, _p(off, c_up), _p(cnt, c_up), C.byref(total)))
        return off, cnt, total.value

    def index_build(self, off, cnt, v_begin=0, v_end=None):
        v_end = self.n if v_end is None else v_end
        if v_end <= v_begin:
            return np.empty(0, np.int32)
        lo = int(off[v_begin])
        hi = int(off[v_end - 1] + cnt[v_end - 1])
        dest = np.empty(hi - lo, np.int32)
        self._ck(self.L.fora_index_build(self.h, _p(off, c_up), _p(cnt, c_up), v_begin, v_end, _p(dest, c_ip)))
        return dest

    def index_build_stat(self):
        """(walks, hops, walk-kernel ms) of the last index_build call"""
        w, h, ms = C.c_uint64(0), C.c_uint64(0), C.c_double(0)
        self._ck(self.L.fora_index_build_stat(self.h, C.byref(w), C.byref(h), C.byref(ms)))
        return w.value, h.value, ms.value

    def index_upload(self, off, cnt, dest):
        off = np.ascontiguousarray(off, np.uint64)
        cnt = np.ascontiguousarray(cnt, np.uint64)
        dest = np.ascontiguousarray(dest, np.int32)
        self._ck(self.L.fora_index_upload(self.h, _p(off, c_up), _p(cnt, c_up), _p(dest, c_ip), len(dest)))

    def power_iteration(self, source, iters=100):
        ppr = np.empty(self.n)
        self._ck(self.L.fora_power_iteration(self.h, source, iters, _p(ppr, c_dp)))
        return ppr


class Group:
    """Several GPUs of one box answering ONE whole-graph query together (fora_group_*, include/fora_b200.h): the library itself
    issues the NCCL broadcast of the compacted push state and the all-reduce of the dense vectors."""

    def __init__(self, n_gpus, seed=1, devices=None):
        self.L = lib()
        self.h = C.c_void_p()
        dev = None if devices is None else np.ascontiguousarray(devices, np.int32)
        rc = self.L.fora_group_create(n_gpus, _p(dev, c_ip), seed, C.byref(self.h))
        if rc:
            raise ForaError("fora_group_create failed: %s" % self.L.fora_group_last_error(None).decode())
        self.engines = []
        for i in range(n_gpus):  # views of the group's contexts (owned by the group)
            E = Engine.__new__(Engine)
            E.L, E.h, E.n, E.m_decl, E.params = self.L, C.c_void_p(self.L.fora_group_ctx(self.h, i)), 0, 0, Params()
            E.close = lambda: None
            self.engines.append(E)

    def upload_graph(self, n, m_decl, out_ptr, out_col):
        for E in self.engines:
            E.upload_graph(n, m_decl, out_ptr, out_col)

    def configure(self, *a, **kw):
        r = None
        for E in self.engines:
            r = E.configure(*a, **kw)
        return r

    def query_split(self, source, query_id=0, want_ppr=True):
        n = self.engines[0].n
        ppr = np.empty(n) if want_ppr else None
        st, tm = QueryStat(), SplitTiming()
        rc = self.L.fora_group_query_split(self.h, int(source), int(query_id), _p(ppr, c_dp), C.byref(st), C.byref(tm))
        if rc:
            raise ForaError("rc=%d: %s" % (rc, self.L.fora_group_last_error(self.h).decode()))
        return ppr, st.as_dict(), tm.as_dict()

    def close(self):
        if getattr(self, "h", None):
            for E in self.engines:
                E.h = None
            self.L.fora_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
