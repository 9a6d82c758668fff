/* include/fora_b200.h -- C ABI of the B200-native FORA engine (libfora_b200.so).
 *
 * The reference (wangsibovictor/fora) has no plugin/FFI interface: its query path is a set of
 * free functions communicating through process-wide globals.  Each entry
 * point below replaces one of those call sites; the reference file:line it stands in for is
 * cited on the declaration.  Plain pointers and sizes only, every call returns 0 on success or
 * a negative FORA_E* code (text via fora_last_error); nothing here runs on the CPU except the
 * explicitly host-side helpers (loader, parameter derivation, archive I/O), and the library
 * refuses to work without a CUDA device -- there is no CPU fallback.
 *
 * Threading: one fora_ctx per host thread / per GPU; calls on one ctx must be serialised.
 */
#ifndef FORA_B200_H
#define FORA_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fora_ctx fora_ctx;

enum {
    FORA_OK = 0,
    FORA_EINVAL = -1,  /* bad argument / call order */
    FORA_ECUDA = -2,   /* CUDA runtime error (no device, launch failure, out of memory) */
    FORA_EIO = -3,     /* file missing / malformed */
    FORA_ERANGE = -4   /* id >= n in the edge list (the reference asserts, graph.h:155-156) */
};

/* --algo values (config.h:41-45) */
enum { FORA_ALGO_FORA = 0, FORA_ALGO_FWDPUSH = 1, FORA_ALGO_MC = 2, FORA_ALGO_BIPPR = 3 };

/* ------------------------------------------------------------------------------------------
 * Host-side helpers (no GPU work).  Same arithmetic, expression order and file formats as the
 * reference so that doubles and layouts are bit-identical.
 * ---------------------------------------------------------------------------------------- */
/* Graph::init_nm, graph.h:48-64 */
int fora_host_read_attribute(const char* path, int32_t* n, int64_t* m);
/* Graph::init_graph edge scan, graph.h:152-160: returns kept edges (self loops dropped, duplicates and
 * file order kept) or a negative code; src/dst may be NULL to count only. */
int64_t fora_host_read_edges(const char* path, int32_t n, int32_t* src, int32_t* dst);
/* g[u].push_back(v) / gr[v].push_back(u), graph.h:158-159, as CSR: ptr int64[n+1], col int32[kept] */
int fora_host_csr_from_edges(int32_t n, int64_t n_edges, const int32_t* src, const int32_t* dst,
                             int64_t* out_ptr, int32_t* out_col, int64_t* in_ptr, int32_t* in_col);
/* deterministic synthetic power-law graph of a named shape; returns kept edges */
int64_t fora_host_synth_edges(int32_t n, int64_t m, uint64_t seed, double exponent, double dangling_frac,
                              int32_t* src, int32_t* dst);

/* algo.h:442-496 parameter derivation; which: 0 fora_setting, 1 fora_topk_setting,
 * 2 montecarlo_setting, 3 bippr_setting, 4 fwdpush_setting.  Inputs delta/pfail as set by
 * init_parameter (graph.h:177-178) or the top-k drivers. */
int fora_host_setting(int which, int32_t n, int64_t m, double epsilon, double delta, double pfail, double alpha,
                      int opt, double rmax_scale, double* rmax, double* omega);

/* ------------------------------------------------------------------------------------------
 * Context
 * ---------------------------------------------------------------------------------------- */
/* main() start, fora.cpp:56.  device = CUDA ordinal; seed keys the Philox streams (the reference
 * seeds from time(0), algo.h:107,116). */
int fora_ctx_create(int device, uint64_t seed, fora_ctx** out);
void fora_ctx_destroy(fora_ctx* ctx);
const char* fora_last_error(fora_ctx* ctx); /* ctx may be NULL: last creation error */
/* Run all work on this cudaStream_t (e.g. torch's current stream); NULL = the ctx's own stream. */
int fora_ctx_set_stream(fora_ctx* ctx, void* cuda_stream);
/* number of query slots processed concurrently per launch (dense state = 16*n bytes per slot) */
int fora_ctx_set_slots(fora_ctx* ctx, int slots);
int fora_ctx_sync(fora_ctx* ctx);
/* The Philox streams of a query are keyed by (ctx seed, GLOBAL query index).  The reference processes its query file in one
 * loop (query.h:1429-1511); a host that shards that list over GPUs, or issues it in several batch calls, passes the list
 * index of the first query of the next fora_query_batch* / fora_topk_batch call here so that no two queries of a run share a
 * random stream and results do not depend on how the list was cut.  Sticky; 0 after fora_ctx_create. */
int fora_ctx_set_query_base(fora_ctx* ctx, uint64_t first_query_index);
/* Opt-in: the queries of one wave (fora_ctx_set_slots of them at a time) draw their random walks from ONE pool.  Walk j from
 * vertex v is computed once per wave and serves every query of the wave that needs more than j walks from v -- what the
 * reference does for ALL queries when it runs --with_idx (query.h:290-307), here with a pool sized by the wave's own needs,
 * rebuilt with fresh randomness for every wave and never stored.  Every query's estimate keeps its distribution and FORA's
 * guarantee; the estimates of queries of the same wave are no longer independent of each other.  Applies to the plain query
 * path (fora_query_batch* with algo fora, no index, one GPU per query); 0 (default) = every query walks for itself. */
int fora_ctx_set_shared_walks(fora_ctx* ctx, int on);

/* ------------------------------------------------------------------------------------------
 * Graph  (Graph graph(folder), fora.cpp:176-177 -> graph.h:37-46,89-163)
 * ---------------------------------------------------------------------------------------- */
/* Host arrays are borrowed for the call only.  in_ptr/in_col may be NULL (needed by BiPPR only).
 * m_decl is attribute.txt's m (used by every formula even when self loops were dropped). */
int fora_graph_upload(fora_ctx* ctx, int32_t n, int64_t m_decl, const int64_t* out_ptr, const int32_t* out_col,
                      const int64_t* in_ptr, const int32_t* in_col);
/* The same construction on the device, from the edge list in file order (graph.h:152-160: self loops dropped,
 * duplicates and file order kept == stable sort by source / by target).  src/dst host int32[n_edges]. */
int fora_graph_build_from_edges(fora_ctx* ctx, int32_t n, int64_t m_decl, const int32_t* src, const int32_t* dst,
                                int64_t n_edges, int with_in);
int fora_graph_download_csr(fora_ctx* ctx, int64_t* out_ptr, int32_t* out_col, int64_t* in_ptr, int32_t* in_col);
int64_t fora_graph_num_edges(fora_ctx* ctx);

/* ------------------------------------------------------------------------------------------
 * Parameters  (globals `config`, config.h:86-138; set by *_setting, algo.h:442-496)
 * ---------------------------------------------------------------------------------------- */
typedef struct fora_params {
    double alpha, epsilon, delta, pfail, rmax, omega, rmax_scale;
    int32_t opt, balanced, with_idx;
    uint32_t k;
    /* --balanced: the reference compares wall-clock push seconds with omega*rsum*(1-alpha)*4e-7
     * (query.h:821-839,867-877).  Here the same loop is driven by a device cost model:
     *   push cost += cost_edge*edges + cost_vertex*vertices + cost_level*levels   per round
     *   walk cost  = omega*rsum*(1-alpha)*cost_walk   (cost_walk/140 when the index suffices)
     * 0 in all four selects the built-in B200 calibration. */
    double cost_walk, cost_edge, cost_vertex, cost_level;
} fora_params;
int fora_params_set(fora_ctx* ctx, const fora_params* p);
int fora_params_get(fora_ctx* ctx, fora_params* p);

/* ------------------------------------------------------------------------------------------
 * Per-query statistics (what the reference accumulates in Timer ids 3/5/6 and the counters
 * num_total_rw / num_hit_idx, algo.h:39-40, plus what the roofline accounting needs)
 * ---------------------------------------------------------------------------------------- */
typedef struct fora_query_stat {
    double rsum;        /* sum of residues after push (== the reference's rsum up to rounding) */
    double final_rmax;  /* r_max the push stopped at (query.h:878-880) */
    uint64_t n_walks;   /* num_total_rw */
    uint64_t n_idx_hits;/* num_hit_idx */
    uint64_t walk_hops; /* neighbour fetches */
    uint64_t edges_pushed, vertices_pushed, push_levels, push_rounds;
    uint64_t n_sources; /* vertices with non-zero residue */
} fora_query_stat;

typedef struct fora_batch_timing { /* GPU milliseconds, CUDA events on the work stream */
    float total_ms, push_ms, walk_ms, plan_ms, topk_ms, copy_ms; /* phases (push_ms includes seeding + host round trips) */
    uint64_t kernel_launches;
    /* the two hot kernels alone, summed over their launches (roofline accounting) */
    float push_kernel_ms, walk_kernel_ms;
    uint64_t push_kernel_launches, walk_kernel_launches;
} fora_batch_timing;

/* ------------------------------------------------------------------------------------------
 * Push (test hooks == the reference's push functions)
 * ---------------------------------------------------------------------------------------- */
/* forward_local_update_linear(s, graph, rsum, rmax), algo.h:954-1018: fresh state, source pushed
 * unconditionally.  reserve/residue: host double[n] (may be NULL). */
int fora_push_only(fora_ctx* ctx, int32_t source, double rmax, double* reserve, double* residue, double* rsum,
                   fora_query_stat* stat);
/* forward_local_update_linear_topk, algo.h:1020-1093: begin installs residue[s]=1 (query.h:849-856),
 * each round continues at a (smaller) rmax from the state the previous round left. */
int fora_push_begin(fora_ctx* ctx, int32_t source);
int fora_push_round(fora_ctx* ctx, double rmax, double* reserve, double* residue, double* rsum, fora_query_stat* stat);
/* reverse_local_update_linear(t, graph), algo.h:703-751 (needs the in-CSR) */
int fora_reverse_push(fora_ctx* ctx, int32_t target, double rmax, double* reserve, double* residue);

/* ------------------------------------------------------------------------------------------
 * Walks (test hook == random_walk / random_walk_no_zero_hop, algo.h:124-166)
 * ---------------------------------------------------------------------------------------- */
int fora_random_walks(fora_ctx* ctx, int32_t start, int64_t count, int no_zero_hop, int32_t* dest, uint64_t* hops);
/* residue-seeded walk phase on a caller-provided push state: compute_ppr_with_fwdidx{,_opt},
 * query.h:255-413.  reserve/residue host double[n]; ppr host double[n]. */
int fora_compute_ppr(fora_ctx* ctx, const double* reserve, const double* residue, double rsum, double* ppr,
                     fora_query_stat* stat);

/* ------------------------------------------------------------------------------------------
 * Queries
 * ---------------------------------------------------------------------------------------- */
/* The per-query loop of query(), query.h:1429-1511: fora_query_basic (query.h:841),
 * montecarlo_query (query.h:16), bippr_query (query.h:71), fwdpush (query.h:1503-1508).
 * sources: host int32[n_q].  ppr: host double[n_q*n] or NULL (the reference discards the vector).
 * stats: host fora_query_stat[n_q] or NULL.  timing may be NULL. */
int fora_query_batch(fora_ctx* ctx, int algo, const int32_t* sources, int32_t n_q, double* ppr,
                     fora_query_stat* stats, fora_batch_timing* timing);
/* The same queries with a COMPACTED result: for query i the (node id, value) pairs of every entry >= threshold (unordered) at
 * ids/values[offsets[i] .. offsets[i+1]).  The reference never ships the vector anywhere (query.h:1471-1476); FORA's guarantee
 * covers pi >= delta = 1/n only (algo.h:455-463), so threshold = 1/n keeps everything the guarantee speaks about at a few per
 * cent of the dense vector's bytes.  ids/values: host buffers of cap_total entries (pinned memory makes the copies overlap the
 * next wave); offsets: host uint64[n_q+1].  FORA_ERANGE when a query yields more than cap_per_query entries or the batch more
 * than cap_total. */
int fora_query_batch_sparse(fora_ctx* ctx, int algo, const int32_t* sources, int32_t n_q, double threshold, uint64_t cap_per_query,
                            uint64_t cap_total, int32_t* ids, double* values, uint64_t* offsets, fora_query_stat* stats,
                            fora_batch_timing* timing);
/* Same work with inputs/outputs resident in HBM: sources device int32[n_q]; the PPR vectors stay
 * on the device (last `slots` queries readable through fora_device_ppr). */
int fora_query_batch_device(fora_ctx* ctx, int algo, const int32_t* d_sources, int32_t n_q,
                            fora_query_stat* stats, fora_batch_timing* timing);
/* device pointer (double[n], internal vertex order -- see fora_device_to_original) of the PPR vector of slot `slot`
 * after the last batch */
void* fora_device_ppr(fora_ctx* ctx, int slot);

/* ---- one query split over several GPUs (whole-graph SSPPR at Twitter scale) ----
 * The reference computes `ppr = reserve + sum of walk increments` on one core (query.h:255-413); here every
 * GPU receives the same push state (device pointers of slot 0 below, filled e.g. by an NCCL broadcast), walks
 * chunk range `part` of `nparts` of the SAME walk plan (Philox keys do not depend on the split) into its own dense
 * vector, and the caller reduces the vectors (ncclAllReduce / torch.distributed over NVLink).  part 0 starts from
 * the reserve, the others from zero, so the sum equals the single-GPU result up to fp64 summation order. */
int fora_prepare_slots(fora_ctx* ctx); /* allocate the per-slot state so the pointers below exist */
/* NOTE: the engine renumbers vertices internally (descending in-degree, for L2 locality).  Everything that crosses the
 * ABI in HOST buffers is in ORIGINAL ids; the raw device pointers below expose the internal order (identical on every
 * GPU that holds the same graph, so they can be broadcast / reduced as they are) -- use fora_device_to_original to
 * obtain a vector indexed by original id. */
void* fora_device_reserve(fora_ctx* ctx, int slot); /* device double[n], internal order */
void* fora_device_residue(fora_ctx* ctx, int slot); /* device double[n], internal order */
int fora_device_to_original(fora_ctx* ctx, const double* d_internal, double* d_original); /* device -> device */
int fora_compute_ppr_part_device(fora_ctx* ctx, double rsum, uint32_t query_id, uint32_t part, uint32_t nparts,
                                 fora_query_stat* stat); /* result in fora_device_reserve(ctx, 0) */

/* ---- the same split inside ONE process: a group of GPUs, NCCL collectives issued by the library (loaded at run time) ----
 * fora_group_create opens one context per GPU (same seed) and one NCCL communicator per GPU; upload the graph and set the
 * parameters on every fora_group_ctx(g, i).  fora_group_query_split then answers ONE whole-graph query with all GPUs: GPU 0
 * pushes (forward_local_update_linear_topk rounds of fora_query_basic, query.h:848-884 -- with G GPUs a walk costs 1/G, so
 * --balanced stops the push earlier), the compacted (vertex, residue) list is broadcast (ncclBroadcast), every GPU walks chunk
 * range g of G of the same walk plan (compute_ppr_with_fwdidx{,_opt}, query.h:255-413), one ncclAllReduce sums the dense
 * vectors.  ppr: host double[n] (original ids) or NULL.  The result equals the single-GPU result up to fp64 summation order. */
typedef struct fora_group fora_group;
typedef struct fora_split_timing { /* GPU milliseconds (CUDA events); walk_ms = slowest GPU */
    float total_ms, push_ms, bcast_ms, walk_ms, reduce_ms;
    uint32_t n_gpus;
    uint64_t bcast_bytes, reduce_bytes;
} fora_split_timing;
int fora_group_create(int n_gpus, const int* devices /* NULL: 0..n_gpus-1 */, uint64_t seed, fora_group** out);
void fora_group_destroy(fora_group* g);
int fora_group_size(fora_group* g);
fora_ctx* fora_group_ctx(fora_group* g, int i);
const char* fora_group_last_error(fora_group* g); /* g may be NULL: last creation error */
int fora_group_query_split(fora_group* g, int32_t source, uint32_t query_id, double* ppr, fora_query_stat* stat, fora_split_timing* timing);

/* get_topk(), query.h:1139-1190: k (node,value) pairs per query, descending, unfilled = (0,0.0)
 * (algo.h:592-610).  iters: per-query refinement rounds (num_iter_topk) or NULL. */
int fora_topk_batch(fora_ctx* ctx, int algo, const int32_t* sources, int32_t n_q, uint32_t k, int32_t* nodes,
                    double* values, int32_t* iters, fora_query_stat* stats, fora_batch_timing* timing);
/* topk_ppr() on a caller-provided dense vector (algo.h:592-610) */
int fora_topk_of(fora_ctx* ctx, const double* ppr, uint32_t k, int32_t* nodes, double* values);

/* ------------------------------------------------------------------------------------------
 * Walk index  (build(), build.h:302-366; deserialize_idx(), build.h:194-207)
 * ---------------------------------------------------------------------------------------- */
/* pass 1, build.h:325-334: offsets/counts are a pure function of degrees, rmax, omega, alpha, opt */
int fora_index_info(fora_ctx* ctx, uint64_t* offsets, uint64_t* counts, uint64_t* total);
/* pass 2, build.h:344-354, for sources [v_begin, v_end): dest receives the slice
 * [offsets[v_begin], offsets[v_end]) (host int32).  Sharding by source range = multi-GPU build. */
int fora_index_build(fora_ctx* ctx, const uint64_t* offsets, const uint64_t* counts, int32_t v_begin, int32_t v_end,
                     int32_t* dest);
/* walks, neighbour hops and walk-kernel milliseconds (CUDA events) of the last fora_index_build call on this ctx (the reference
 * prints only wall-clock build time, build.h:360-363; these feed the roofline line of the index build) */
int fora_index_build_stat(fora_ctx* ctx, uint64_t* walks, uint64_t* hops, double* kernel_ms);
int fora_index_upload(fora_ctx* ctx, const uint64_t* offsets, const uint64_t* counts, const int32_t* dest, uint64_t len);

/* ------------------------------------------------------------------------------------------
 * Ground truth: fwd_power_iteration, query.h:1192-1224 (dense, on the device)
 * ---------------------------------------------------------------------------------------- */
int fora_power_iteration(fora_ctx* ctx, int32_t source, int iters, double* ppr);

#ifdef __cplusplus
}
#endif
#endif
