/* oracle/fora_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's (wangsibovictor/fora) query path, used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER for the CUDA
 * path.  Nothing in fora_b200/ may include, link or call this.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function here against
 * the unmodified reference compiled by oracle/Makefile (oracle/_ref/libfora_ref.so, Boost
 * replaced by oracle/boost_shim) and against the fixtures under tests/golden/ generated from
 * that build (tests/golden/make_golden.py).  The reference ships no tests or golden vectors of
 * its own.
 *
 * Conventions: ids int32, edge offsets int64, all real arithmetic in double with the
 * reference's expression order (compile with -ffp-contract=off).  Each function cites the
 * reference file:line it restates.
 */
#ifndef FORA_ORACLE_H
#define FORA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_state orc_state;

/* ---- graph (graph.h:48-64, 89-163) ---- */
int orc_read_attribute(const char* path, int* n, long long* m);
/* returns number of edges kept (self loops dropped, duplicates kept, file order), -1 on error,
 * -2 if an id >= n (reference asserts).  Pass src=dst=NULL to count only. */
long long orc_read_edges(const char* path, int n, int* src, int* dst);
void orc_csr_from_edges(int n, long long n_edges, const int* src, const int* dst,
                        long long* out_ptr, int* out_col, long long* in_ptr, int* in_col);

/* ---- parameters (graph.h:173-183, algo.h:442-496) ---- */
void orc_fora_setting(long long m, double epsilon, double delta, double pfail, double alpha, int opt,
                      double rmax_scale, double* rmax, double* omega);
void orc_fora_topk_setting(long long m, double epsilon, double delta, double pfail, double rmax_scale,
                           double* rmax, double* omega);
void orc_montecarlo_setting(double epsilon, double delta, double pfail, double* omega);
void orc_bippr_setting(long long m, double epsilon, double delta, double pfail, double rmax_scale,
                       double* rmax, double* omega);
void orc_fwdpush_setting(int n, long long m, double epsilon, double delta, double rmax_scale, double* rmax);

/* ---- state ---- */
/* Arrays are borrowed (caller keeps them alive).  m_decl is attribute.txt's m. */
orc_state* orc_create(int n, long long m_decl, const long long* out_ptr, const int* out_col,
                      const long long* in_ptr, const int* in_col);
void orc_destroy(orc_state* st);
void orc_seed(orc_state* st, uint64_t seed);
/* alpha, epsilon, delta, pfail, rmax, omega, rmax_scale; flags opt/balanced/with_idx; k */
void orc_set_params(orc_state* st, double alpha, double epsilon, double delta, double pfail, double rmax,
                    double omega, double rmax_scale, int opt, int balanced, int with_idx, unsigned k);
/* nil sentinels as set by query() (-1, query.h:1464) or topk() (-9, query.h:1344); mode 0 = query
 * (ppr.init_keys: dense), 1 = topk (ppr sparse) */
void orc_init_state(orc_state* st, double nil, int topk_mode);

/* dense copies (0 where absent) + insertion-ordered key lists */
void orc_get_fwd(orc_state* st, double* reserve, double* residue);
int orc_get_residue_occur(orc_state* st, int* keys);
int orc_get_reserve_occur(orc_state* st, int* keys);
void orc_get_bwd(orc_state* st, double* reserve, double* residue);
void orc_get_ppr(orc_state* st, double* ppr);
void orc_set_fwd(orc_state* st, const double* reserve, const double* residue); /* install a push state */
void orc_get_counters(orc_state* st, unsigned long long* out8);
/* out8: total_rw, hit_idx, walk_hops, edges_pushed, vertices_pushed, push_levels, topk_iters, rounds */
void orc_reset_counters(orc_state* st);

/* ---- push ---- */
/* forward_local_update_linear, FIFO (algo.h:954-1018) */
double orc_forward_push_fifo(orc_state* st, int s, double rmax, double init_residual);
/* forward_local_update_linear_topk (algo.h:1020-1093) */
void orc_push_topk_begin(orc_state* st, int s);
double orc_push_topk_round(orc_state* st, int s, double rmax, double lowest_rmax);
int orc_push_topk_candidates(orc_state* st, int* out);
/* Frontier-synchronous schedule of the SAME push rule (the CUDA path's schedule):
 * seed_all=0: level 0 = {s} unconditionally (algo.h:973); seed_all=1: level 0 = every v with
 * residue/d_out >= rmax (resumable round).  fresh=1 clears state and sets residue[s]=init. */
double orc_forward_push_sync(orc_state* st, int s, double rmax, int fresh, int seed_all);
/* reverse_local_update_linear (algo.h:703-751); sync=1 uses the frontier-synchronous schedule
 * with the intended "skip below rmax" semantics instead of the FIFO with its early break */
void orc_reverse_push(orc_state* st, int t, double rmax, double init_residual, int sync);

/* ---- walks (algo.h:124-166) ---- */
int orc_random_walk(orc_state* st, int start);
int orc_random_walk_no_zero_hop(orc_state* st, int start);
void orc_random_walks(orc_state* st, int start, long long count, int no_zero_hop, int* dest);

/* ---- residue-seeded Monte Carlo (query.h:243-413, 521-750) ---- */
void orc_compute_ppr_with_reserve(orc_state* st);
void orc_compute_ppr_with_fwdidx(orc_state* st, double rsum);
void orc_compute_ppr_with_fwdidx_opt(orc_state* st, double rsum);
void orc_compute_ppr_with_fwdidx_topk(orc_state* st, double rsum);
void orc_compute_ppr_with_fwdidx_topk_with_bound(orc_state* st, double rsum);
/* per-source walk plan of compute_ppr_with_fwdidx{,_opt}: for every key of residue in order:
 * count[i], incre[i]; returns number of keys.  Pure arithmetic (query.h:270,314-317 / 349,400-404). */
long long orc_walk_plan(orc_state* st, double rsum, int opt, int* keys, unsigned long long* counts, double* incre);

/* ---- query drivers ---- */
/* fora_query_basic (query.h:841-907).  balanced_mode: 0 = wall clock as the reference,
 * 1 = deterministic cost model  used = c_edge*edges + c_vertex*vertices + c_level*levels  with the
 * frontier-synchronous push (the CUDA path's restatement).  Returns rsum after push. */
double orc_fora_query_basic(orc_state* st, int s, int balanced_mode, int sync_push, double walk_cost,
                            double c_edge, double c_vertex, double c_level, double* final_rmax);
void orc_montecarlo_query(orc_state* st, int s, int topk_variant);              /* query.h:16-69 */
void orc_bippr_query(orc_state* st, int s, int topk_variant, int sync_push);    /* query.h:71-193 */
void orc_fwdpush_query(orc_state* st, int s);                                   /* query.h:1503-1508 */
void orc_fora_query_topk_new(orc_state* st, int s, int sync_push);              /* query.h:972-1045 */
void orc_fora_query_topk_with_bound(orc_state* st, int s, int sync_push);       /* query.h:909-969 */

/* ---- top-k (algo.h:578-610) and precision (algo.h:524-572) ---- */
double orc_kth_ppr(orc_state* st, unsigned k);
double orc_topk_ppr(orc_state* st, unsigned k, int* nodes, double* values);
void orc_precision(unsigned k, int n_est, const int* est_nodes, const double* est_values, int n_exact,
                   const int* exact_nodes, const double* exact_values, double* precision, double* recall);

/* ---- index (build.h:302-366) ---- */
/* offsets/counts only: pure function of degrees, rmax, omega, alpha, opt (build.h:325-334) */
unsigned long long orc_index_info(orc_state* st, unsigned long long* offsets, unsigned long long* counts);
void orc_index_build(orc_state* st, const unsigned long long* offsets, const unsigned long long* counts, int* dest);
void orc_index_set(orc_state* st, const unsigned long long* offsets, const unsigned long long* counts, const int* dest);

/* ---- ground truth (query.h:1192-1224), dense, `iters` sweeps ---- */
void orc_power_iteration(orc_state* st, int s, int iters, double* ppr);

#ifdef __cplusplus
}
#endif
#endif
