/* oracle/fora_oracle.c -- TEST INFRASTRUCTURE ONLY (see fora_oracle.h for the contract).
 *
 * Plain-C restatement of the reference query path of wangsibovictor/fora.  Written from the
 * reference's behaviour, function by function; each block cites the reference file:line it
 * follows.  Parity is pinned against the unmodified reference (oracle/_ref) by
 * tests/test_oracle_vs_reference.py and tests/golden/.
 */
#define _POSIX_C_SOURCE 200809L
#include "fora_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------
 * iMap<T> (mylib.h:278-425): dense value array + list of keys in first-insert order;
 * exist(p) <=> data[p] != nil.  Kept faithful (including duplicate keys in `occ` when a value
 * returns to nil and is inserted again) because iteration order / duplicates are observable
 * in bippr_query (query.h:101-112).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    double* d;
    int* occ;
    long nocc, cap;
    double nil;
    int n;
} imapd;

static void imap_init(imapd* m, int n, double nil) { /* initialize(), mylib.h:302-314 */
    m->n = n;
    m->nil = nil;
    free(m->d);
    m->d = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) m->d[i] = nil;
    if (m->cap < n || !m->occ) {
        free(m->occ);
        m->cap = n > 20 ? n : 20;
        m->occ = (int*)malloc(sizeof(int) * (size_t)m->cap);
    }
    m->nocc = 0;
}
static void imap_push_occ(imapd* m, int p) {
    if (m->nocc == m->cap) {
        m->cap *= 2;
        m->occ = (int*)realloc(m->occ, sizeof(int) * (size_t)m->cap);
    }
    m->occ[m->nocc++] = p;
}
static void imap_init_keys(imapd* m, int n, double nil) { /* init_keys(), mylib.h:326-339 */
    imap_init(m, n, nil);
    for (int i = 0; i < n; ++i) {
        m->d[i] = 0;
        imap_push_occ(m, i);
    }
}
static void imap_clean(imapd* m) { /* clean(), mylib.h:315-323 */
    for (long i = 0; i < m->nocc; ++i) m->d[m->occ[i]] = m->nil;
    m->nocc = 0;
}
static inline int imap_exist(const imapd* m, int p) { return !(m->d[p] == m->nil); }
static inline void imap_insert(imapd* m, int p, double v) { /* insert(), mylib.h:387-399 */
    if (m->d[p] == m->nil) imap_push_occ(m, p);
    m->d[p] = v;
}
static void imap_reset(imapd* m, double v) { /* reset_zero_values / reset_one_values */
    for (int i = 0; i < m->n; ++i) m->d[i] = v;
}
static void imap_free(imapd* m) {
    free(m->d);
    free(m->occ);
    memset(m, 0, sizeof *m);
}
static int cmp_int(const void* a, const void* b) {
    int x = *(const int*)a, y = *(const int*)b;
    return (x > y) - (x < y);
}
static void imap_sort_occ(imapd* m) { qsort(m->occ, (size_t)m->nocc, sizeof(int), cmp_int); } /* iVector::Sort */

/* ------------------------------------------------------------------------------------------ */
struct orc_state {
    int n;
    long long m;
    const long long *out_ptr, *in_ptr;
    const int *out_col, *in_col;
    /* config (config.h:86-138) */
    double alpha, epsilon, delta, pfail, rmax, omega, rmax_scale;
    int opt, balanced, with_idx;
    unsigned k;
    /* global working state (algo.h:28-49) */
    imapd reserve, residue; /* fwd_idx.first / .second */
    imapd breserve, bresidue;
    imapd ppr, rw_counter, upper_bounds, lower_bounds, topk_filter;
    int topk_mode;
    double zero_ppr_upper_bound, threshold;
    unsigned long long total_rw, hit_idx, walk_hops, edges_pushed, vertices_pushed, push_levels, topk_iters, rounds;
    /* index */
    const unsigned long long *idx_off, *idx_cnt;
    const int* idx_dest;
    /* resumable push worklist */
    int* forward_from;
    long n_forward, cap_forward;
    double rsum;
    /* function-local statics of the reference, latched at first use */
    int latched_basic, latched_new, latched_bound, latched_bounds_fn, latched_stop;
    double basic_lowest, new_init_delta, new_pfail, new_lowest, bound_pfail, bound_lowest, sb_min_ppr, sb_sqrt_min_ppr,
        stop_error;
    /* rng */
    uint64_t s[4];
};

static inline int deg_out(const orc_state* st, int v) { return (int)(st->out_ptr[v + 1] - st->out_ptr[v]); }

/* ---- rng: xoshiro256++ seeded by splitmix64 (the reference's streams are seeded from time(0)
 * and are not reproducible, algo.h:107,116; any good generator is an equivalent restatement) */
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t next64(orc_state* st) {
    uint64_t* s = st->s;
    const uint64_t r = rotl(s[0] + s[3], 23) + s[0];
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return r;
}
void orc_seed(orc_state* st, uint64_t seed) {
    for (int i = 0; i < 4; ++i) {
        uint64_t z = (seed += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        st->s[i] = z ^ (z >> 31);
    }
}
static inline int drand_stop(orc_state* st) { /* drand(), algo.h:114-119: Bernoulli(alpha) */
    return (double)(next64(st) >> 11) * (1.0 / 9007199254740992.0) <= st->alpha;
}
static inline unsigned long lrand32(orc_state* st) { return (unsigned long)(next64(st) >> 32); } /* lrand(), algo.h:105 */

/* ------------------------------------------------------------------------------------------
 * graph (graph.h:48-64 init_nm, graph.h:152-160 init_graph)
 * ---------------------------------------------------------------------------------------- */
int orc_read_attribute(const char* path, int* n, long long* m) {
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int c;
    while ((c = fgetc(f)) != EOF && c != '=') {}
    if (fscanf(f, "%d", n) != 1) { fclose(f); return -1; }
    while ((c = fgetc(f)) != EOF && c != '=') {}
    if (fscanf(f, "%lld", m) != 1) { fclose(f); return -1; }
    fclose(f);
    return 0;
}

long long orc_read_edges(const char* path, int n, int* src, int* dst) {
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int t1, t2;
    long long k = 0;
    while (fscanf(f, "%d%d", &t1, &t2) == 2) {
        if (!(t1 < n) || !(t2 < n)) { fclose(f); return -2; } /* graph.h:155-156 asserts */
        if (t1 == t2) continue;                                 /* graph.h:157 */
        if (src) { src[k] = t1; dst[k] = t2; }
        ++k;
    }
    fclose(f);
    return k;
}

/* g[t1].push_back(t2); gr[t2].push_back(t1) in file order (graph.h:158-159) == stable counting sort */
void orc_csr_from_edges(int n, long long ne, const int* src, const int* dst, long long* out_ptr, int* out_col,
                        long long* in_ptr, int* in_col) {
    for (int i = 0; i <= n; ++i) out_ptr[i] = in_ptr[i] = 0;
    for (long long e = 0; e < ne; ++e) {
        if (src[e] == dst[e]) continue;
        out_ptr[src[e] + 1]++;
        in_ptr[dst[e] + 1]++;
    }
    for (int i = 0; i < n; ++i) { out_ptr[i + 1] += out_ptr[i]; in_ptr[i + 1] += in_ptr[i]; }
    long long* po = (long long*)malloc(sizeof(long long) * (size_t)(n + 1));
    long long* pi = (long long*)malloc(sizeof(long long) * (size_t)(n + 1));
    memcpy(po, out_ptr, sizeof(long long) * (size_t)(n + 1));
    memcpy(pi, in_ptr, sizeof(long long) * (size_t)(n + 1));
    for (long long e = 0; e < ne; ++e) {
        if (src[e] == dst[e]) continue;
        out_col[po[src[e]]++] = dst[e];
        in_col[pi[dst[e]]++] = src[e];
    }
    free(po);
    free(pi);
}

/* ------------------------------------------------------------------------------------------
 * parameters
 * ---------------------------------------------------------------------------------------- */
void orc_fora_setting(long long m, double epsilon, double delta, double pfail, double alpha, int opt,
                      double rmax_scale, double* rmax, double* omega) { /* algo.h:455-463 */
    double r = epsilon * sqrt(delta / 3 / m / log(2 / pfail));
    if (opt) r *= rmax_scale / (1 - alpha);
    else r *= rmax_scale;
    *rmax = r;
    *omega = (2 + epsilon) * log(2 / pfail) / delta / epsilon / epsilon;
}
void orc_fora_topk_setting(long long m, double epsilon, double delta, double pfail, double rmax_scale,
                           double* rmax, double* omega) { /* algo.h:466-474 (both branches identical) */
    double r = epsilon * sqrt(delta / 3 / m / log(2 / pfail));
    r *= sqrt(1.0 * m * r) * rmax_scale * 3;
    *rmax = r;
    *omega = (2 + epsilon) * log(2 / pfail) / delta / epsilon / epsilon;
}
void orc_montecarlo_setting(double epsilon, double delta, double pfail, double* omega) { /* algo.h:477-483 */
    *omega = 3 * log(2 / pfail) / epsilon / epsilon / delta;
}
void orc_bippr_setting(long long m, double epsilon, double delta, double pfail, double rmax_scale, double* rmax,
                       double* omega) { /* algo.h:442-447 */
    double r = epsilon * sqrt(m * 1.0 * delta / 3.0 / log(2.0 / pfail));
    r *= rmax_scale;
    *rmax = r;
    *omega = r * 3 * log(2.0 / pfail) / delta / epsilon / epsilon;
}
void orc_fwdpush_setting(int n, long long m, double epsilon, double delta, double rmax_scale, double* rmax) {
    *rmax = rmax_scale * delta * epsilon * n / m; /* algo.h:495 */
}

/* ------------------------------------------------------------------------------------------ */
orc_state* orc_create(int n, long long m_decl, const long long* out_ptr, const int* out_col, const long long* in_ptr,
                      const int* in_col) {
    orc_state* st = (orc_state*)calloc(1, sizeof *st);
    st->n = n;
    st->m = m_decl;
    st->out_ptr = out_ptr; st->out_col = out_col; st->in_ptr = in_ptr; st->in_col = in_col;
    st->alpha = 0.2;               /* config.h:27,132 */
    st->delta = st->pfail = 1.0 / n; /* init_parameter, graph.h:177-178 */
    st->rmax_scale = 1;
    st->k = 500;
    st->zero_ppr_upper_bound = 1.0;
    orc_seed(st, 1);
    orc_init_state(st, -1, 0);
    return st;
}
void orc_destroy(orc_state* st) {
    if (!st) return;
    imap_free(&st->reserve); imap_free(&st->residue); imap_free(&st->breserve); imap_free(&st->bresidue);
    imap_free(&st->ppr); imap_free(&st->rw_counter); imap_free(&st->upper_bounds); imap_free(&st->lower_bounds);
    imap_free(&st->topk_filter);
    free(st->forward_from);
    free(st);
}
void orc_set_params(orc_state* st, double alpha, double epsilon, double delta, double pfail, double rmax, double omega,
                    double rmax_scale, int opt, int balanced, int with_idx, unsigned k) {
    st->alpha = alpha; st->epsilon = epsilon; st->delta = delta; st->pfail = pfail; st->rmax = rmax;
    st->omega = omega; st->rmax_scale = rmax_scale; st->opt = opt; st->balanced = balanced;
    st->with_idx = with_idx; st->k = k;
}
void orc_init_state(orc_state* st, double nil, int topk_mode) {
    int n = st->n;
    st->topk_mode = topk_mode;
    imap_init(&st->reserve, n, nil);
    imap_init(&st->residue, n, nil);
    imap_init(&st->breserve, n, 0); /* bwd_idx nil stays the zero-initialised global (query.h:1434-1435) */
    imap_init(&st->bresidue, n, 0);
    if (topk_mode) { /* topk(), query.h:1343-1357 */
        imap_init_keys(&st->rw_counter, n, nil);
        imap_init_keys(&st->upper_bounds, n, nil);
        imap_init_keys(&st->lower_bounds, n, nil);
        imap_init(&st->ppr, n, nil);
        imap_init(&st->topk_filter, n, nil);
    } else { /* query(), query.h:1427,1437 */
        imap_init_keys(&st->ppr, n, 0);
        imap_init(&st->rw_counter, n, 0);
    }
}
void orc_get_fwd(orc_state* st, double* reserve, double* residue) {
    for (int i = 0; i < st->n; ++i) {
        if (reserve) reserve[i] = imap_exist(&st->reserve, i) ? st->reserve.d[i] : 0.0;
        if (residue) residue[i] = imap_exist(&st->residue, i) ? st->residue.d[i] : 0.0;
    }
}
int orc_get_residue_occur(orc_state* st, int* keys) {
    if (keys) memcpy(keys, st->residue.occ, sizeof(int) * (size_t)st->residue.nocc);
    return (int)st->residue.nocc;
}
int orc_get_reserve_occur(orc_state* st, int* keys) {
    if (keys) memcpy(keys, st->reserve.occ, sizeof(int) * (size_t)st->reserve.nocc);
    return (int)st->reserve.nocc;
}
void orc_get_bwd(orc_state* st, double* reserve, double* residue) {
    for (int i = 0; i < st->n; ++i) {
        if (reserve) reserve[i] = imap_exist(&st->breserve, i) ? st->breserve.d[i] : 0.0;
        if (residue) residue[i] = imap_exist(&st->bresidue, i) ? st->bresidue.d[i] : 0.0;
    }
}
void orc_get_ppr(orc_state* st, double* ppr) {
    for (int i = 0; i < st->n; ++i) ppr[i] = imap_exist(&st->ppr, i) ? st->ppr.d[i] : 0.0;
}
void orc_set_fwd(orc_state* st, const double* reserve, const double* residue) {
    imap_clean(&st->reserve);
    imap_clean(&st->residue);
    for (int i = 0; i < st->n; ++i) {
        if (reserve[i] != 0.0) imap_insert(&st->reserve, i, reserve[i]);
        if (residue[i] != 0.0) imap_insert(&st->residue, i, residue[i]);
    }
}
void orc_get_counters(orc_state* st, unsigned long long* o) {
    o[0] = st->total_rw; o[1] = st->hit_idx; o[2] = st->walk_hops; o[3] = st->edges_pushed;
    o[4] = st->vertices_pushed; o[5] = st->push_levels; o[6] = st->topk_iters; o[7] = st->rounds;
}
void orc_reset_counters(orc_state* st) {
    st->total_rw = st->hit_idx = st->walk_hops = st->edges_pushed = st->vertices_pushed = st->push_levels =
        st->topk_iters = st->rounds = 0;
}

/* ------------------------------------------------------------------------------------------
 * forward push, FIFO (algo.h:954-1018)
 * ---------------------------------------------------------------------------------------- */
double orc_forward_push_fifo(orc_state* st, int s, double rmax, double init_residual) {
    const double alpha = st->alpha;
    double rsum = 1.0;
    imap_clean(&st->reserve);
    imap_clean(&st->residue);
    unsigned char* idx = (unsigned char*)calloc((size_t)st->n, 1);
    if (deg_out(st, s) == 0) { /* algo.h:961-965 */
        imap_insert(&st->reserve, s, 1);
        free(idx);
        return 0;
    }
    const double myeps = rmax;
    long cap = st->n > 16 ? st->n : 16, qn = 0, left = 0;
    int* q = (int*)malloc(sizeof(int) * (size_t)cap);
    q[qn++] = s;
    imap_insert(&st->residue, s, init_residual);
    idx[s] = 1;
    while (left < qn) {
        int v = q[left];
        idx[v] = 0;
        left++;
        double v_residue = st->residue.d[v];
        st->residue.d[v] = 0;
        if (!imap_exist(&st->reserve, v)) imap_insert(&st->reserve, v, v_residue * alpha);
        else st->reserve.d[v] += v_residue * alpha;
        int out_neighbor = deg_out(st, v);
        rsum -= v_residue * alpha;
        st->vertices_pushed++;
        if (out_neighbor == 0) { /* algo.h:993-1000: dangling mass returns to the source */
            st->residue.d[s] += v_residue * (1 - alpha);
            if (deg_out(st, s) > 0 && st->residue.d[s] / deg_out(st, s) >= myeps && idx[s] != 1) {
                idx[s] = 1;
                if (qn == cap) { cap *= 2; q = (int*)realloc(q, sizeof(int) * (size_t)cap); }
                q[qn++] = s;
            }
            continue;
        }
        double avg_push_residual = ((1.0 - alpha) * v_residue) / out_neighbor;
        st->edges_pushed += (unsigned long long)out_neighbor;
        for (long long e = st->out_ptr[v]; e < st->out_ptr[v + 1]; ++e) {
            int next = st->out_col[e];
            if (!imap_exist(&st->residue, next)) imap_insert(&st->residue, next, avg_push_residual);
            else st->residue.d[next] += avg_push_residual;
            if (st->residue.d[next] / deg_out(st, next) >= myeps && idx[next] != 1) {
                idx[next] = 1;
                if (qn == cap) { cap *= 2; q = (int*)realloc(q, sizeof(int) * (size_t)cap); }
                q[qn++] = next;
            }
        }
    }
    free(q);
    free(idx);
    st->rsum = rsum;
    return rsum;
}

/* ------------------------------------------------------------------------------------------
 * resumable push (algo.h:1020-1093)
 * ---------------------------------------------------------------------------------------- */
static void ff_push(orc_state* st, int** arr, long* n, long* cap, int v) {
    (void)st;
    if (*n == *cap) {
        *cap = *cap ? *cap * 2 : 1024;
        *arr = (int*)realloc(*arr, sizeof(int) * (size_t)*cap);
    }
    (*arr)[(*n)++] = v;
}
void orc_push_topk_begin(orc_state* st, int s) { /* e.g. query.h:849-856 */
    st->n_forward = 0;
    ff_push(st, &st->forward_from, &st->n_forward, &st->cap_forward, s);
    imap_clean(&st->reserve);
    imap_clean(&st->residue);
    st->rsum = 1.0;
    imap_insert(&st->residue, s, st->rsum);
}
double orc_push_topk_round(orc_state* st, int s, double rmax, double lowest_rmax) {
    const double alpha = st->alpha, myeps = rmax;
    double rsum = st->rsum;
    unsigned char* in_forward = (unsigned char*)calloc((size_t)st->n, 1);
    unsigned char* in_next = (unsigned char*)calloc((size_t)st->n, 1);
    int* next_from = NULL;
    long n_next = 0, cap_next = 0;
    for (long i = 0; i < st->n_forward; ++i) in_forward[st->forward_from[i]] = 1;
    long i = 0;
    while (i < st->n_forward) {
        int v = st->forward_from[i];
        i++;
        in_forward[v] = 0;
        if (st->residue.d[v] / deg_out(st, v) >= myeps) {
            int out_neighbor = deg_out(st, v);
            double v_residue = st->residue.d[v];
            st->residue.d[v] = 0;
            if (!imap_exist(&st->reserve, v)) imap_insert(&st->reserve, v, v_residue * alpha);
            else st->reserve.d[v] += v_residue * alpha;
            rsum -= v_residue * alpha;
            st->vertices_pushed++;
            if (out_neighbor == 0) { /* algo.h:1051-1064 */
                st->residue.d[s] += v_residue * (1 - alpha);
                if (deg_out(st, s) > 0 && in_forward[s] != 1 && st->residue.d[s] / deg_out(st, s) >= myeps) {
                    ff_push(st, &st->forward_from, &st->n_forward, &st->cap_forward, s);
                    in_forward[s] = 1;
                } else if (deg_out(st, s) >= 0 && in_next[s] != 1 && st->residue.d[s] / deg_out(st, s) >= lowest_rmax) {
                    ff_push(st, &next_from, &n_next, &cap_next, s);
                    in_next[s] = 1;
                }
                continue;
            }
            double avg_push_residual = ((1 - alpha) * v_residue) / out_neighbor;
            st->edges_pushed += (unsigned long long)out_neighbor;
            for (long long e = st->out_ptr[v]; e < st->out_ptr[v + 1]; ++e) {
                int next = st->out_col[e];
                if (!imap_exist(&st->residue, next)) imap_insert(&st->residue, next, avg_push_residual);
                else st->residue.d[next] += avg_push_residual;
                if (in_forward[next] != 1 && st->residue.d[next] / deg_out(st, next) >= myeps) {
                    ff_push(st, &st->forward_from, &st->n_forward, &st->cap_forward, next);
                    in_forward[next] = 1;
                } else if (in_next[next] != 1 && st->residue.d[next] / deg_out(st, next) >= lowest_rmax) {
                    ff_push(st, &next_from, &n_next, &cap_next, next);
                    in_next[next] = 1;
                }
            }
        } else if (in_next[v] != 1 && st->residue.d[v] / deg_out(st, v) >= lowest_rmax) { /* algo.h:1084-1089 */
            ff_push(st, &next_from, &n_next, &cap_next, v);
            in_next[v] = 1;
        }
    }
    free(st->forward_from); /* forward_from = next_forward_from, algo.h:1092 */
    st->forward_from = next_from;
    st->n_forward = n_next;
    st->cap_forward = cap_next;
    free(in_forward);
    free(in_next);
    st->rsum = rsum;
    st->rounds++;
    return rsum;
}
int orc_push_topk_candidates(orc_state* st, int* out) {
    if (out) memcpy(out, st->forward_from, sizeof(int) * (size_t)st->n_forward);
    return (int)st->n_forward;
}

/* ------------------------------------------------------------------------------------------
 * Frontier-synchronous schedule of the same push rule.  This is NOT in the reference: it is
 * the schedule of the CUDA kernel (DESIGN.md "push"), restated here so that the GPU result can
 * be checked to 1e-9 (only the order of floating-point additions into one residue differs).
 * Per-vertex arithmetic is the reference's (algo.h:984-1008): reserve += alpha*r,
 * neighbours += ((1-alpha)*r)/d_out, dangling mass -> source, threshold residue/d_out >= rmax.
 *   level k:  A) every frontier vertex reads and zeroes its residue;  B) all scatters land;
 *   a vertex joins level k+1 when a scatter moves residue/d_out from < rmax to >= rmax.
 * ---------------------------------------------------------------------------------------- */
static inline int crosses(double oldv, double newv, int d, double rmax) {
    return !(oldv / d >= rmax) && (newv / d >= rmax);
}
double orc_forward_push_sync(orc_state* st, int s, double rmax, int fresh, int seed_all) {
    const double alpha = st->alpha;
    const int n = st->n;
    if (fresh) {
        imap_clean(&st->reserve);
        imap_clean(&st->residue);
        st->rsum = 1.0;
        if (deg_out(st, s) == 0) {
            imap_insert(&st->reserve, s, 1);
            st->rsum = 0;
            return 0;
        }
        imap_insert(&st->residue, s, 1.0);
    }
    int* cur = (int*)malloc(sizeof(int) * (size_t)n);
    int* nxt = (int*)malloc(sizeof(int) * (size_t)n);
    double* rv = (double*)malloc(sizeof(double) * (size_t)n);
    long ncur = 0, nnxt = 0;
    if (seed_all) {
        for (int v = 0; v < n; ++v)
            if (imap_exist(&st->residue, v) && st->residue.d[v] / deg_out(st, v) >= rmax) cur[ncur++] = v;
    } else {
        cur[ncur++] = s;
    }
    double rsum = st->rsum;
    while (ncur > 0) {
        st->push_levels++;
        for (long i = 0; i < ncur; ++i) { /* phase A */
            int v = cur[i];
            double r = st->residue.d[v];
            st->residue.d[v] = 0;
            if (!imap_exist(&st->reserve, v)) imap_insert(&st->reserve, v, r * alpha);
            else st->reserve.d[v] += r * alpha;
            rsum -= r * alpha;
            rv[i] = r;
            st->vertices_pushed++;
        }
        nnxt = 0;
        for (long i = 0; i < ncur; ++i) { /* phase B */
            int v = cur[i];
            int d = deg_out(st, v);
            if (d == 0) {
                double inc = rv[i] * (1 - alpha);
                double o = imap_exist(&st->residue, s) ? st->residue.d[s] : 0.0;
                imap_insert(&st->residue, s, o + inc);
                if (crosses(o, o + inc, deg_out(st, s), rmax)) nxt[nnxt++] = s;
                continue;
            }
            double inc = ((1.0 - alpha) * rv[i]) / d;
            st->edges_pushed += (unsigned long long)d;
            for (long long e = st->out_ptr[v]; e < st->out_ptr[v + 1]; ++e) {
                int u = st->out_col[e];
                double o = imap_exist(&st->residue, u) ? st->residue.d[u] : 0.0;
                imap_insert(&st->residue, u, o + inc);
                if (crosses(o, o + inc, deg_out(st, u), rmax)) nxt[nnxt++] = u;
            }
        }
        int* t = cur; cur = nxt; nxt = t;
        ncur = nnxt;
    }
    free(cur); free(nxt); free(rv);
    st->rsum = rsum;
    st->rounds++;
    return rsum;
}

/* ------------------------------------------------------------------------------------------
 * backward push (algo.h:703-751)
 * ---------------------------------------------------------------------------------------- */
void orc_reverse_push(orc_state* st, int t, double rmax, double init_residual, int sync) {
    const double alpha = st->alpha, myeps = rmax;
    imap_clean(&st->breserve);
    imap_clean(&st->bresidue);
    const int n = st->n;
    if (!sync) {
        unsigned char* idx = (unsigned char*)calloc((size_t)n, 1);
        long cap = 64, qn = 0, left = 0;
        int* q = (int*)malloc(sizeof(int) * (size_t)cap);
        q[qn++] = t;
        imap_insert(&st->bresidue, t, init_residual);
        idx[t] = 1;
        while (left < qn) {
            int v = q[left];
            idx[v] = 0;
            left++;
            if (st->bresidue.d[v] < myeps) break; /* algo.h:725-726: break, not continue */
            if (!imap_exist(&st->breserve, v)) imap_insert(&st->breserve, v, st->bresidue.d[v] * alpha);
            else st->breserve.d[v] += st->bresidue.d[v] * alpha;
            double residual = (1 - alpha) * st->bresidue.d[v];
            st->bresidue.d[v] = 0;
            for (long long e = st->in_ptr[v]; e < st->in_ptr[v + 1]; ++e) {
                int next = st->in_col[e];
                int cnt = deg_out(st, next);
                if (!imap_exist(&st->bresidue, next)) imap_insert(&st->bresidue, next, residual / cnt);
                else st->bresidue.d[next] += residual / cnt;
                if (st->bresidue.d[next] > myeps && idx[next] != 1) {
                    idx[next] = 1;
                    if (qn == cap) { cap *= 2; q = (int*)realloc(q, sizeof(int) * (size_t)cap); }
                    q[qn++] = next;
                }
            }
        }
        free(q);
        free(idx);
        return;
    }
    /* frontier-synchronous restatement (CUDA schedule): level 0 = {t} if init >= rmax (algo.h:725);
     * a vertex joins the next level when a scatter moves its residue from <= rmax to > rmax (algo.h:743) */
    int* cur = (int*)malloc(sizeof(int) * (size_t)n);
    int* nxt = (int*)malloc(sizeof(int) * (size_t)n);
    double* rv = (double*)malloc(sizeof(double) * (size_t)n);
    long ncur = 0, nnxt = 0;
    imap_insert(&st->bresidue, t, init_residual);
    if (!(init_residual < myeps)) cur[ncur++] = t;
    while (ncur > 0) {
        for (long i = 0; i < ncur; ++i) {
            int v = cur[i];
            double r = st->bresidue.d[v];
            st->bresidue.d[v] = 0;
            if (!imap_exist(&st->breserve, v)) imap_insert(&st->breserve, v, r * alpha);
            else st->breserve.d[v] += r * alpha;
            rv[i] = r;
        }
        nnxt = 0;
        for (long i = 0; i < ncur; ++i) {
            int v = cur[i];
            double residual = (1 - alpha) * rv[i];
            for (long long e = st->in_ptr[v]; e < st->in_ptr[v + 1]; ++e) {
                int u = st->in_col[e];
                double inc = residual / deg_out(st, u);
                double o = st->bresidue.d[u]; /* nil == 0 here */
                imap_insert(&st->bresidue, u, o + inc);
                if (!(o > myeps) && (o + inc > myeps)) nxt[nnxt++] = u;
            }
        }
        int* tt = cur; cur = nxt; nxt = tt;
        ncur = nnxt;
    }
    free(cur); free(nxt); free(rv);
}

/* ------------------------------------------------------------------------------------------
 * random walks (algo.h:124-166)
 * ---------------------------------------------------------------------------------------- */
int orc_random_walk(orc_state* st, int start) { /* algo.h:124-142 */
    int cur = start;
    if (deg_out(st, start) == 0) return start;
    while (1) {
        if (drand_stop(st)) return cur;
        int d = deg_out(st, cur);
        if (d) {
            unsigned long k = lrand32(st) % (unsigned long)d;
            cur = st->out_col[st->out_ptr[cur] + (long long)k];
            st->walk_hops++;
        } else {
            cur = start;
        }
    }
}
int orc_random_walk_no_zero_hop(orc_state* st, int start) { /* algo.h:144-166 */
    int cur = start;
    if (deg_out(st, start) == 0) return start;
    unsigned long k = lrand32(st) % (unsigned long)deg_out(st, cur);
    cur = st->out_col[st->out_ptr[cur] + (long long)k];
    st->walk_hops++;
    while (1) {
        if (drand_stop(st)) return cur;
        int d = deg_out(st, cur);
        if (d) {
            k = lrand32(st) % (unsigned long)d;
            cur = st->out_col[st->out_ptr[cur] + (long long)k];
            st->walk_hops++;
        } else {
            cur = start;
        }
    }
}
void orc_random_walks(orc_state* st, int start, long long count, int no_zero_hop, int* dest) {
    for (long long i = 0; i < count; ++i)
        dest[i] = no_zero_hop ? orc_random_walk_no_zero_hop(st, start) : orc_random_walk(st, start);
}

/* ------------------------------------------------------------------------------------------
 * ppr = reserve + residue-seeded walks
 * ---------------------------------------------------------------------------------------- */
static inline void ppr_add(orc_state* st, int des, double inc) { /* "if(!ppr.exist(des)) insert else +=" */
    if (!imap_exist(&st->ppr, des)) imap_insert(&st->ppr, des, inc);
    else st->ppr.d[des] += inc;
}

void orc_compute_ppr_with_reserve(orc_state* st) { /* query.h:243-253 */
    imap_clean(&st->ppr);
    for (long i = 0; i < st->reserve.nocc; ++i) {
        int node_id = st->reserve.occ[i];
        double reserve = st->reserve.d[node_id];
        if (reserve) imap_insert(&st->ppr, node_id, reserve);
    }
}

long long orc_walk_plan(orc_state* st, double check_rsum, int opt, int* keys, unsigned long long* counts,
                        double* incre) {
    if (check_rsum == 0.0) return 0;
    if (opt) check_rsum *= (1 - st->alpha);
    unsigned long long num_random_walk = (unsigned long long)(st->omega * check_rsum);
    long long k = 0;
    for (long i = 0; i < st->residue.nocc; ++i) {
        int source = st->residue.occ[i];
        if (opt && !imap_exist(&st->residue, source)) continue;
        double residual = opt ? st->residue.d[source] * (1 - st->alpha) : st->residue.d[source];
        unsigned long num_s_rw = (unsigned long)ceil(residual / check_rsum * num_random_walk);
        double a_s = residual / check_rsum * num_random_walk / num_s_rw;
        double ppr_incre = a_s * check_rsum / num_random_walk;
        keys[k] = source; counts[k] = num_s_rw; incre[k] = ppr_incre;
        ++k;
    }
    return k;
}

static void ppr_from_fwd(orc_state* st, double check_rsum, int opt) { /* query.h:255-327 / 334-413 */
    imap_reset(&st->ppr, 0.0); /* ppr.reset_zero_values() */
    for (long i = 0; i < st->reserve.nocc; ++i) {
        int node_id = st->reserve.occ[i];
        st->ppr.d[node_id] = st->reserve.d[node_id];
    }
    if (check_rsum == 0.0) return;
    if (opt) check_rsum *= (1 - st->alpha);
    unsigned long long num_random_walk = (unsigned long long)(st->omega * check_rsum);
    if (st->with_idx) imap_sort_occ(&st->residue); /* query.h:278 */
    for (long i = 0; i < st->residue.nocc; ++i) {
        int source = st->residue.occ[i];
        double residual;
        if (opt) {
            if (!imap_exist(&st->residue, source)) continue;
            st->ppr.d[source] += st->residue.d[source] * st->alpha;
            residual = st->residue.d[source] * (1 - st->alpha);
        } else {
            residual = st->residue.d[source];
        }
        unsigned long num_s_rw = (unsigned long)ceil(residual / check_rsum * num_random_walk);
        double a_s = residual / check_rsum * num_random_walk / num_s_rw;
        double ppr_incre = a_s * check_rsum / num_random_walk;
        st->total_rw += num_s_rw;
        unsigned long from_idx = 0;
        if (st->with_idx) { /* query.h:290-307: always from the start of the source's slice */
            unsigned long have = (unsigned long)st->idx_cnt[source];
            from_idx = num_s_rw > have ? have : num_s_rw;
            for (unsigned long k = 0; k < from_idx; ++k) {
                int des = st->idx_dest[st->idx_off[source] + k];
                st->ppr.d[des] += ppr_incre;
            }
            st->hit_idx += from_idx;
        }
        for (unsigned long j = from_idx; j < num_s_rw; ++j) {
            int des = opt ? orc_random_walk_no_zero_hop(st, source) : orc_random_walk(st, source);
            st->ppr.d[des] += ppr_incre;
        }
    }
}
void orc_compute_ppr_with_fwdidx(orc_state* st, double rsum) { ppr_from_fwd(st, rsum, 0); }
void orc_compute_ppr_with_fwdidx_opt(orc_state* st, double rsum) { ppr_from_fwd(st, rsum, 1); }

void orc_compute_ppr_with_fwdidx_topk(orc_state* st, double check_rsum) { /* query.h:521-636 */
    orc_compute_ppr_with_reserve(st);
    if (check_rsum == 0.0) return;
    check_rsum *= (1 - st->alpha);
    if (st->with_idx) {
        imap_sort_occ(&st->residue);
        for (long i = 0; i < st->residue.nocc; ++i) {
            int source = st->residue.occ[i];
            double residual = st->residue.d[source];
            ppr_add(st, source, residual * st->alpha);
            residual *= (1 - st->alpha);
            unsigned long num_s_rw = (unsigned long)ceil(residual * st->omega);
            double a_s = residual * st->omega / num_s_rw;
            double ppr_incre = a_s / st->omega;
            st->total_rw += num_s_rw;
            unsigned long num_used_idx = (unsigned long)st->rw_counter.d[source];
            unsigned long num_remaining_idx = (unsigned long)st->idx_cnt[source] - num_used_idx;
            unsigned long take = num_s_rw <= num_remaining_idx ? num_s_rw : num_remaining_idx;
            for (unsigned long k = 0; k < take; ++k) ppr_add(st, st->idx_dest[st->idx_off[source] + num_used_idx + k], ppr_incre);
            st->rw_counter.d[source] = (double)(num_used_idx + take);
            st->hit_idx += take;
            for (unsigned long j = take; j < num_s_rw; ++j) ppr_add(st, orc_random_walk_no_zero_hop(st, source), ppr_incre);
        }
    } else { /* query.h:615-632: plain random_walk, no (1-alpha) scaling, no alpha*r credit */
        for (long i = 0; i < st->residue.nocc; ++i) {
            int source = st->residue.occ[i];
            double residual = st->residue.d[source];
            unsigned long num_s_rw = (unsigned long)ceil(residual * st->omega);
            double a_s = residual * st->omega / num_s_rw;
            double ppr_incre = a_s / st->omega;
            st->total_rw += num_s_rw;
            for (unsigned long j = 0; j < num_s_rw; ++j) ppr_add(st, orc_random_walk(st, source), ppr_incre);
        }
    }
}

static double calculate_lambda(double rsum, double pfail, double upper_bound, long total_rw_num) { /* algo.h:1169-1174 */
    return 1.0 / 3 * log(2 / pfail) * rsum / total_rw_num +
           sqrt(4.0 / 9.0 * log(2.0 / pfail) * log(2.0 / pfail) * rsum * rsum +
                8 * total_rw_num * log(2.0 / pfail) * rsum * upper_bound) /
               2.0 / total_rw_num;
}

static void set_ppr_bounds(orc_state* st, double rsum, long total_rw_num) { /* algo.h:1178-1261 */
    if (!st->latched_bounds_fn) {
        st->sb_min_ppr = 1.0 / st->n;
        st->sb_sqrt_min_ppr = sqrt(1.0 / st->n);
        st->latched_bounds_fn = 1;
    }
    const double min_ppr = st->sb_min_ppr, sqrt_min_ppr = st->sb_sqrt_min_ppr;
    double epsilon_v_div = sqrt(2.67 * rsum * log(2.0 / st->pfail) / total_rw_num);
    double default_epsilon_v = epsilon_v_div / sqrt_min_ppr;
    st->zero_ppr_upper_bound = calculate_lambda(rsum, st->pfail, st->zero_ppr_upper_bound, total_rw_num);
    for (long i = 0; i < st->ppr.nocc; ++i) {
        int nodeid = st->ppr.occ[i];
        if (st->ppr.d[nodeid] <= 0) continue;
        double reserve = 0.0;
        if (imap_exist(&st->reserve, nodeid)) reserve = st->reserve.d[nodeid];
        double epsilon_a = 1.0;
        if (imap_exist(&st->upper_bounds, nodeid)) {
            if (st->upper_bounds.d[nodeid] > reserve)
                epsilon_a = calculate_lambda(rsum, st->pfail, st->upper_bounds.d[nodeid] - reserve, total_rw_num);
            else
                epsilon_a = calculate_lambda(rsum, st->pfail, 1 - reserve, total_rw_num);
        } else {
            epsilon_a = calculate_lambda(rsum, st->pfail, 1.0 - reserve, total_rw_num);
        }
        double ub_eps_a = st->ppr.d[nodeid] + epsilon_a;
        double lb_eps_a = st->ppr.d[nodeid] - epsilon_a;
        if (!(lb_eps_a > 0)) lb_eps_a = 0;
        double epsilon_v = default_epsilon_v;
        if (imap_exist(&st->reserve, nodeid) && st->reserve.d[nodeid] > min_ppr) {
            if (imap_exist(&st->lower_bounds, nodeid)) reserve = fmax(reserve, st->lower_bounds.d[nodeid]);
            epsilon_v = epsilon_v_div / sqrt(reserve);
        } else {
            if (st->lower_bounds.d[nodeid] > 0) epsilon_v = epsilon_v_div / sqrt(st->lower_bounds.d[nodeid]);
        }
        double ub_eps_v = 1.0, lb_eps_v = 0.0;
        if (1.0 - epsilon_v > 0) {
            ub_eps_v = st->ppr.d[nodeid] / (1.0 - epsilon_v);
            lb_eps_v = st->ppr.d[nodeid] / (1.0 + epsilon_v);
        }
        double up_bound = fmin(fmin(ub_eps_a, ub_eps_v), 1.0);
        double low_bound = fmax(fmax(lb_eps_a, lb_eps_v), reserve);
        if (up_bound > 0) imap_insert(&st->upper_bounds, nodeid, up_bound);
        if (low_bound >= 0) imap_insert(&st->lower_bounds, nodeid, low_bound);
    }
}

void orc_compute_ppr_with_fwdidx_topk_with_bound(orc_state* st, double check_rsum) { /* query.h:639-750 */
    orc_compute_ppr_with_reserve(st);
    if (check_rsum == 0.0) return;
    long num_random_walk = (long)(st->omega * check_rsum);
    long real_num_rand_walk = 0;
    if (st->with_idx) {
        imap_sort_occ(&st->residue);
        for (long i = 0; i < st->residue.nocc; ++i) {
            int source = st->residue.occ[i];
            double residual = st->residue.d[source];
            long num_s_rw = (long)ceil(residual * st->omega);
            double a_s = residual / check_rsum * num_random_walk / num_s_rw;
            double ppr_incre = a_s * check_rsum / num_random_walk;
            st->total_rw += (unsigned long long)num_s_rw;
            real_num_rand_walk += num_s_rw;
            long num_used_idx = 0;
            int source_cnt_exist = imap_exist(&st->rw_counter, source);
            if (source_cnt_exist) num_used_idx = (long)st->rw_counter.d[source];
            long num_remaining_idx = (long)st->idx_cnt[source] - num_used_idx;
            /* query.h:678,698: destinations are read from the start of the slice (cursor not applied) */
            long take = num_s_rw <= num_remaining_idx ? num_s_rw : num_remaining_idx;
            if (take < 0) take = 0;
            for (long k = 0; k < take; ++k) ppr_add(st, st->idx_dest[st->idx_off[source] + (unsigned long long)k], ppr_incre);
            if (source_cnt_exist) st->rw_counter.d[source] += (double)take;
            else imap_insert(&st->rw_counter, source, (double)take);
            st->hit_idx += (unsigned long long)take;
            for (long j = 0; j < num_s_rw - take; ++j) ppr_add(st, orc_random_walk(st, source), ppr_incre);
        }
    } else {
        for (long i = 0; i < st->residue.nocc; ++i) {
            int source = st->residue.occ[i];
            double residual = st->residue.d[source];
            long num_s_rw = (long)ceil(residual / check_rsum * num_random_walk);
            double a_s = residual / check_rsum * num_random_walk / num_s_rw;
            real_num_rand_walk += num_s_rw;
            double ppr_incre = a_s * check_rsum / num_random_walk;
            for (long j = 0; j < num_s_rw; ++j) ppr_add(st, orc_random_walk(st, source), ppr_incre);
        }
    }
    if (st->delta < st->threshold) set_ppr_bounds(st, check_rsum, real_num_rand_walk);
    else st->zero_ppr_upper_bound = calculate_lambda(check_rsum, st->pfail, st->zero_ppr_upper_bound, real_num_rand_walk);
}

/* ------------------------------------------------------------------------------------------
 * top-k selection (algo.h:578-610)
 * ---------------------------------------------------------------------------------------- */
typedef struct { int node; double val; } nv;
static int cmp_nv_desc(const void* a, const void* b) {
    const nv *x = (const nv*)a, *y = (const nv*)b;
    if (x->val > y->val) return -1;
    if (x->val < y->val) return 1;
    return (x->node > y->node) - (x->node < y->node);
}
static int cmp_d_desc(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x < y) - (x > y);
}
double orc_kth_ppr(orc_state* st, unsigned k) { /* algo.h:578-590 (loop index taken as 0) */
    long cnt = st->ppr.nocc;
    if (cnt < (long)k || k == 0) return 0.0;
    double* t = (double*)malloc(sizeof(double) * (size_t)cnt);
    for (long i = 0; i < cnt; ++i) t[i] = st->ppr.d[st->ppr.occ[i]];
    qsort(t, (size_t)cnt, sizeof(double), cmp_d_desc);
    double r = t[k - 1];
    free(t);
    return r;
}
double orc_topk_ppr(orc_state* st, unsigned k, int* nodes, double* values) { /* algo.h:592-610 */
    /* the reference copies into an unordered_map first: one entry per distinct key */
    unsigned char* seen = (unsigned char*)calloc((size_t)st->n, 1);
    nv* t = (nv*)malloc(sizeof(nv) * (size_t)(st->ppr.nocc > 0 ? st->ppr.nocc : 1));
    long cnt = 0;
    for (long i = 0; i < st->ppr.nocc; ++i) {
        int v = st->ppr.occ[i];
        if (seen[v]) continue;
        seen[v] = 1;
        t[cnt].node = v;
        t[cnt].val = st->ppr.d[v];
        cnt++;
    }
    qsort(t, (size_t)cnt, sizeof(nv), cmp_nv_desc);
    for (unsigned i = 0; i < k; ++i) {
        if ((long)i < cnt) { nodes[i] = t[i].node; values[i] = t[i].val; }
        else { nodes[i] = 0; values[i] = 0.0; }
    }
    free(t);
    free(seen);
    return values[k - 1];
}

void orc_precision(unsigned k, int n_est, const int* est_nodes, const double* est_values, int n_exact,
                   const int* exact_nodes, const double* exact_values, double* precision, double* recall) {
    /* algo.h:524-572: both ratios use |exact_map| as denominator; entries with value <= 0 ignored */
    int size_e = (int)k < n_exact ? (int)k : n_exact;
    double rec = 0, pre = 0;
    int exact_cnt = 0;
    for (int i = 0; i < size_e; ++i) {
        if (!(exact_values[i] > 0)) continue;
        int dup = 0;
        for (int j = 0; j < i; ++j) if (exact_values[j] > 0 && exact_nodes[j] == exact_nodes[i]) { dup = 1; break; }
        if (dup) continue;
        exact_cnt++;
        for (int j = 0; j < n_est; ++j)
            if (est_values[j] > 0 && est_nodes[j] == exact_nodes[i]) { rec++; break; }
    }
    for (int j = 0; j < n_est; ++j) {
        if (!(est_values[j] > 0)) continue;
        int dup = 0;
        for (int q = 0; q < j; ++q) if (est_values[q] > 0 && est_nodes[q] == est_nodes[j]) { dup = 1; break; }
        if (dup) continue;
        for (int i = 0; i < size_e; ++i)
            if (exact_values[i] > 0 && exact_nodes[i] == est_nodes[j]) { pre++; break; }
    }
    *recall = rec * 1.0 / exact_cnt;
    *precision = pre * 1.0 / exact_cnt;
}

/* ------------------------------------------------------------------------------------------
 * query drivers
 * ---------------------------------------------------------------------------------------- */
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

double orc_fora_query_basic(orc_state* st, int v, int balanced_mode, int sync_push, double walk_cost, double c_edge,
                            double c_vertex, double c_level, double* final_rmax) { /* query.h:841-907 */
    double rsum = 1.0;
    double rmax = st->rmax;
    if (st->balanced) {
        orc_push_topk_begin(st, v);
        if (!st->latched_basic) { /* const static, query.h:858-860 */
            double min_delta = 1.0 / st->n;
            st->basic_lowest = st->opt ? st->epsilon * sqrt(min_delta / 3 / st->m / log(2 / st->pfail)) / (1 - st->alpha)
                                       : st->epsilon * sqrt(min_delta / 3 / st->m / log(2 / st->pfail));
            st->latched_basic = 1;
        }
        double used_time = 0;
        rmax = st->rmax * 8;
        if (deg_out(st, v) > 0) {
            const double rw_time = balanced_mode ? walk_cost : 0.0000004; /* query.h:822 */
            while (1) {
                /* estimated_random_walk_cost, query.h:826-839 */
                double est;
                if (!st->with_idx || rmax >= st->rmax) est = st->omega * rsum * (1 - st->alpha) * rw_time;
                else est = st->omega * rsum * (1 - st->alpha) * (rw_time / 140);
                if (!(est > used_time)) break;
                unsigned long long e0 = st->edges_pushed, v0 = st->vertices_pushed, l0 = st->push_levels;
                double t0 = now_s();
                if (sync_push) {
                    rsum = orc_forward_push_sync(st, v, rmax, 0, 1);
                } else {
                    rsum = orc_push_topk_round(st, v, rmax, st->basic_lowest);
                }
                if (balanced_mode)
                    used_time += c_edge * (double)(st->edges_pushed - e0) + c_vertex * (double)(st->vertices_pushed - v0) +
                                 c_level * (double)(st->push_levels - l0);
                else
                    used_time += now_s() - t0;
                rmax /= 2;
            }
            rmax *= 2;
        } else {
            rsum = sync_push ? orc_forward_push_sync(st, v, st->rmax, 1, 0) : orc_forward_push_fifo(st, v, st->rmax, 1.0);
        }
    } else {
        rsum = sync_push ? orc_forward_push_sync(st, v, st->rmax, 1, 0) : orc_forward_push_fifo(st, v, st->rmax, 1.0);
    }
    if (final_rmax) *final_rmax = rmax;
    if (st->opt) orc_compute_ppr_with_fwdidx_opt(st, rsum);
    else orc_compute_ppr_with_fwdidx(st, rsum);
    return rsum;
}

void orc_montecarlo_query(orc_state* st, int v, int topk_variant) { /* query.h:16-69 */
    imap_clean(&st->rw_counter);
    if (topk_variant) imap_clean(&st->ppr);
    else imap_reset(&st->ppr, 0.0);
    st->total_rw += (unsigned long long)st->omega;
    for (unsigned long i = 0; i < st->omega; i++) {
        int destination = orc_random_walk(st, v);
        if (!imap_exist(&st->rw_counter, destination)) imap_insert(&st->rw_counter, destination, 1);
        else st->rw_counter.d[destination] += 1;
    }
    for (long i = 0; i < st->rw_counter.nocc; ++i) {
        int node_id = st->rw_counter.occ[i];
        if (topk_variant) {
            if (st->rw_counter.occ[i] > 0) /* query.h:66 tests the node id: node 0 is never reported */
                imap_insert(&st->ppr, node_id, st->rw_counter.d[node_id] * 1.0 / st->omega);
        } else {
            st->ppr.d[node_id] = st->rw_counter.d[node_id] * 1.0 / st->omega;
        }
    }
}

void orc_bippr_query(orc_state* st, int v, int topk_variant, int sync_push) { /* query.h:71-193 */
    if (topk_variant) imap_clean(&st->ppr);
    else imap_reset(&st->ppr, 0.0);
    imap_clean(&st->rw_counter);
    st->total_rw += (unsigned long long)st->omega;
    for (unsigned long i = 0; i < st->omega; i++) {
        int destination = orc_random_walk(st, v);
        if (!imap_exist(&st->rw_counter, destination)) imap_insert(&st->rw_counter, destination, 1);
        else st->rw_counter.d[destination] += 1;
    }
    if (st->rmax < 1.0) {
        for (int i = 0; i < st->n; ++i) {
            orc_reverse_push(st, i, st->rmax, 1.0, sync_push);
            if ((!imap_exist(&st->breserve, v) || 0 == st->breserve.d[v]) && 0 == st->bresidue.nocc) continue;
            if (topk_variant) {
                if (imap_exist(&st->breserve, v) && st->breserve.d[v] > 0) imap_insert(&st->ppr, i, st->breserve.d[v]);
            } else {
                st->ppr.d[i] += st->breserve.d[v];
            }
            if (sync_push) {
                /* intended semantics: each vertex with residue counted once */
                unsigned char* seen = NULL;
                for (long j = 0; j < st->bresidue.nocc; ++j) {
                    int nodeid = st->bresidue.occ[j];
                    int dup = 0;
                    for (long q = 0; q < j; ++q) if (st->bresidue.occ[q] == nodeid) { dup = 1; break; }
                    if (dup) continue;
                    double residual = st->bresidue.d[nodeid];
                    double occur = imap_exist(&st->rw_counter, nodeid) ? st->rw_counter.d[nodeid] : 0;
                    if (topk_variant) { if (occur > 0) ppr_add(st, i, occur * residual / st->omega); }
                    else st->ppr.d[i] += occur * 1.0 / st->omega * residual;
                }
                (void)seen;
            } else {
                for (long j = 0; j < st->bresidue.nocc; ++j) {
                    int nodeid = st->bresidue.occ[j];
                    double residual = st->bresidue.d[nodeid];
                    int occur = imap_exist(&st->rw_counter, nodeid) ? (int)st->rw_counter.d[nodeid] : 0;
                    if (topk_variant) { if (occur > 0) ppr_add(st, i, occur * residual / st->omega); }
                    else st->ppr.d[i] += occur * 1.0 / st->omega * residual;
                }
            }
        }
    } else {
        for (long i = 0; i < st->rw_counter.nocc; ++i) {
            int node_id = st->rw_counter.occ[i];
            if (topk_variant) {
                if (st->rw_counter.d[node_id] > 0) imap_insert(&st->ppr, node_id, st->rw_counter.d[node_id] * 1.0 / st->omega);
            } else {
                st->ppr.d[node_id] = st->rw_counter.d[node_id] * 1.0 / st->omega;
            }
        }
    }
}

void orc_fwdpush_query(orc_state* st, int s) { /* query.h:1503-1508 */
    orc_forward_push_fifo(st, s, st->rmax, 1.0);
    orc_compute_ppr_with_reserve(st);
}

void orc_fora_query_topk_new(orc_state* st, int v, int sync_push) { /* query.h:972-1045 */
    const double min_delta = 1.0 / st->n;
    if (st->k == 0) st->k = 500;
    if (!st->latched_new) { /* const static locals, query.h:974-982 */
        st->new_init_delta = 1.0 / st->k / 10;
        st->new_pfail = 1.0 / st->n / st->n;
        st->new_lowest = st->epsilon * sqrt(min_delta / 3 / st->m / log(2 / st->new_pfail));
        st->latched_new = 1;
    }
    st->pfail = st->new_pfail;
    st->delta = st->new_init_delta;
    double rsum = 1.0;
    orc_push_topk_begin(st, v);
    if (st->with_idx) imap_reset(&st->rw_counter, 0.0);
    while (st->delta >= min_delta) {
        orc_fora_topk_setting(st->m, st->epsilon, st->delta, st->pfail, st->rmax_scale, &st->rmax, &st->omega);
        st->topk_iters++;
        if (deg_out(st, v) == 0) {
            rsum = 0.0;
            imap_insert(&st->reserve, v, 1);
            orc_compute_ppr_with_reserve(st);
            return;
        }
        rsum = sync_push ? orc_forward_push_sync(st, v, st->rmax, 0, 1) : orc_push_topk_round(st, v, st->rmax, st->new_lowest);
        orc_compute_ppr_with_fwdidx_topk(st, rsum);
        double kth = orc_kth_ppr(st, st->k);
        if (kth >= (1 + st->epsilon) * st->delta || st->delta <= min_delta) break;
        st->delta = fmax(min_delta, st->delta / 4.0);
    }
}

static int if_stop(orc_state* st) { /* algo.h:1096-1166 */
    if (orc_kth_ppr(st, st->k) >= 2.0 * st->delta) return 1;
    if (st->delta >= st->threshold) return 0;
    if (!st->latched_stop) { st->stop_error = 1.0 + st->epsilon; st->latched_stop = 1; }
    const double error = st->stop_error, error_2 = st->stop_error;
    const unsigned k = st->k;
    imap_clean(&st->topk_filter);
    long cnt = st->lower_bounds.nocc;
    nv* t = (nv*)malloc(sizeof(nv) * (size_t)(cnt > 0 ? cnt : 1));
    for (long i = 0; i < cnt; ++i) { t[i].node = st->lower_bounds.occ[i]; t[i].val = st->lower_bounds.d[t[i].node]; }
    qsort(t, (size_t)cnt, sizeof(nv), cmp_nv_desc);
    int ok = 1;
    for (unsigned i = 0; i < k && ok; ++i) {
        int node = (long)i < cnt ? t[i].node : 0;
        imap_insert(&st->topk_filter, node, 1);
        double ratio = st->upper_bounds.d[node] / st->lower_bounds.d[node];
        if (ratio > error_2) ok = 0;
    }
    if (!ok) { free(t); return 0; }
    double low_bound_k = (long)(k - 1) < cnt ? t[k - 1].val : 0.0;
    free(t);
    if (low_bound_k <= st->delta) return 0;
    for (long i = 0; i < st->upper_bounds.nocc; ++i) {
        int nodeid = st->upper_bounds.occ[i];
        if (imap_exist(&st->topk_filter, nodeid) || st->ppr.d[nodeid] <= 0) continue; /* raw value: nil (-9) counts as <= 0 */
        double upper_temp = st->upper_bounds.d[nodeid], lower_temp = st->lower_bounds.d[nodeid];
        if (upper_temp > low_bound_k * error) {
            if (upper_temp > (1 + st->epsilon) / (1 - st->epsilon) * lower_temp) continue;
            return 0;
        }
    }
    return 1;
}

void orc_fora_query_topk_with_bound(orc_state* st, int v, int sync_push) { /* query.h:909-969 */
    const double min_delta = 1.0 / st->n, init_delta = 1.0 / 4, ppr_decay_alpha = 0.77; /* config.h:123 */
    st->threshold = (1.0 - ppr_decay_alpha) / pow(500, ppr_decay_alpha) / pow(st->n, 1 - ppr_decay_alpha);
    if (!st->latched_bound) {
        st->bound_pfail = 1.0 / st->n / st->n / log(st->n);
        st->bound_lowest = st->epsilon * sqrt(min_delta / 3 / st->m / log(2 / st->bound_pfail));
        st->latched_bound = 1;
    }
    st->pfail = st->bound_pfail;
    st->delta = init_delta;
    double rsum = 1.0;
    orc_push_topk_begin(st, v);
    st->zero_ppr_upper_bound = 1.0;
    if (st->with_idx) imap_reset(&st->rw_counter, 0.0);
    imap_reset(&st->upper_bounds, 1.0);
    imap_reset(&st->lower_bounds, 0.0);
    while (st->delta >= min_delta) {
        orc_fora_setting(st->m, st->epsilon, st->delta, st->pfail, st->alpha, st->opt, st->rmax_scale, &st->rmax, &st->omega);
        st->topk_iters++;
        if (deg_out(st, v) == 0) {
            rsum = 0.0;
            imap_insert(&st->reserve, v, 1);
            orc_compute_ppr_with_reserve(st);
            return;
        }
        rsum = sync_push ? orc_forward_push_sync(st, v, st->rmax, 0, 1) : orc_push_topk_round(st, v, st->rmax, st->bound_lowest);
        orc_compute_ppr_with_fwdidx_topk_with_bound(st, rsum);
        if (if_stop(st) || st->delta <= min_delta) break;
        st->delta = fmax(min_delta, st->delta / 2.0);
    }
}

/* ------------------------------------------------------------------------------------------
 * index (build.h:302-366)
 * ---------------------------------------------------------------------------------------- */
unsigned long long orc_index_info(orc_state* st, unsigned long long* offsets, unsigned long long* counts) {
    unsigned long long tuned = 0;
    for (int source = 0; source < st->n; ++source) { /* build.h:325-334 */
        unsigned long num_rw;
        size_t d = (size_t)deg_out(st, source);
        if (st->opt) num_rw = (unsigned long)ceil(d * st->rmax * (1 - st->alpha) * st->omega);
        else num_rw = (unsigned long)ceil(d * st->rmax * st->omega);
        offsets[source] = tuned;
        counts[source] = num_rw;
        tuned += num_rw;
    }
    return tuned;
}
void orc_index_build(orc_state* st, const unsigned long long* offsets, const unsigned long long* counts, int* dest) {
    for (int source = 0; source < st->n; ++source) /* build.h:344-354 */
        for (unsigned long long i = 0; i < counts[source]; ++i)
            dest[offsets[source] + i] = st->opt ? orc_random_walk_no_zero_hop(st, source) : orc_random_walk(st, source);
}
void orc_index_set(orc_state* st, const unsigned long long* offsets, const unsigned long long* counts, const int* dest) {
    st->idx_off = offsets; st->idx_cnt = counts; st->idx_dest = dest;
}

/* ------------------------------------------------------------------------------------------
 * ground truth: forward power iteration (query.h:1192-1224), dense restatement.
 * The reference iterates an unordered_map snapshot; the dense sweep visits the same entries in
 * id order, so only floating-point summation order differs.
 * ---------------------------------------------------------------------------------------- */
void orc_power_iteration(orc_state* st, int start, int iters, double* ppr) {
    const int n = st->n;
    const double alpha = st->alpha;
    double* cur = (double*)calloc((size_t)n, sizeof(double));
    double* nxt = (double*)calloc((size_t)n, sizeof(double));
    for (int i = 0; i < n; ++i) ppr[i] = 0.0;
    cur[start] = 1.0;
    for (int it = 0; it < iters; ++it) {
        memset(nxt, 0, sizeof(double) * (size_t)n);
        for (int v = 0; v < n; ++v) {
            double r = cur[v];
            if (!(r > 0)) continue;
            ppr[v] += alpha * r;
            int d = deg_out(st, v);
            double remain = (1 - alpha) * r;
            if (d == 0) {
                nxt[start] += remain;
            } else {
                double avg = remain / d;
                for (long long e = st->out_ptr[v]; e < st->out_ptr[v + 1]; ++e) nxt[st->out_col[e]] += avg;
            }
        }
        double* t = cur; cur = nxt; nxt = t;
    }
    free(cur);
    free(nxt);
}
