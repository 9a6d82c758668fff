// oracle/ref_harness.cpp  --  TEST INFRASTRUCTURE, never linked into the product.
//
// Wraps the UNMODIFIED reference (headers #included where they lie under /root/reference,
// compiled by oracle/Makefile against oracle/boost_shim) behind a C ABI so that tests can
// call the reference's own functions on in-memory graphs and read back the global state
// they leave behind (fwd_idx, ppr, rw_idx, ...).  Output: oracle/_ref/libfora_ref.so.
//
// Only this wrapper is ours; every algorithmic line executed is the reference's.
// One loaded instance == one (graph, config): the reference keeps function-local statics
// sized/derived at first call (algo.h:958,1023; query.h:858-860,911-920,974-982), so tests
// load a fresh copy of the .so per configuration.
#define HEAD_INFO
#include "mylib.h"
#include "graph.h"
#include "config.h"
#include "algo.h"
#include "query.h"
#include "build.h"

#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <unistd.h>

static Graph* g_graph = nullptr;
static std::streambuf* g_cout_buf = nullptr;
static std::ostringstream g_sink;

static void mute() {
    if (!g_cout_buf) g_cout_buf = std::cout.rdbuf(g_sink.rdbuf());
    g_sink.str("");
}

extern "C" {

// ---- graph ---------------------------------------------------------------------------
// Load through the reference's own text loader (graph.h:89-163).
int ref_graph_load_dir(const char* folder) {
    mute();
    config.action = QUERY;
    config.graph_location = folder;
    g_graph = new Graph(folder);
    init_parameter(config, *g_graph);
    return g_graph->n;
}

// Build a Graph from adjacency arrays without touching the text loader (for big graphs).
// Lists are filled exactly as given, so pass them in the order the loader would produce.
int ref_graph_from_csr(int n, long long m_decl, const long long* out_ptr, const int* out_col,
                       const long long* in_ptr, const int* in_col) {
    mute();
    char tmpl[] = "/tmp/fora_ref_XXXXXX";
    char* dir = mkdtemp(tmpl);
    if (!dir) return -1;
    std::string d(dir);
    {
        std::ofstream a(d + "/attribute.txt");
        a << "n=" << n << "\nm=" << m_decl << "\n";
    }
    config.action = GEN_SS_QUERY;  // constructor then reads attribute.txt only (graph.h:41-44)
    g_graph = new Graph(d);
    config.action = QUERY;
    unlink((d + "/attribute.txt").c_str());
    rmdir(d.c_str());
    g_graph->g.assign(n, vector<int>());
    g_graph->gr.assign(n, vector<int>());
    for (int v = 0; v < n; ++v) {
        g_graph->g[v].assign(out_col + out_ptr[v], out_col + out_ptr[v + 1]);
        g_graph->gr[v].assign(in_col + in_ptr[v], in_col + in_ptr[v + 1]);
    }
    init_parameter(config, *g_graph);
    return n;
}

long long ref_graph_m() { return g_graph->m; }
long long ref_graph_num_out_edges() {
    long long s = 0;
    for (auto& l : g_graph->g) s += (long long)l.size();
    return s;
}
void ref_graph_dump(long long* out_ptr, int* out_col, long long* in_ptr, int* in_col) {
    long long o = 0, i = 0;
    for (int v = 0; v < g_graph->n; ++v) {
        out_ptr[v] = o;
        for (int u : g_graph->g[v]) out_col[o++] = u;
        in_ptr[v] = i;
        for (int u : g_graph->gr[v]) in_col[i++] = u;
    }
    out_ptr[g_graph->n] = o;
    in_ptr[g_graph->n] = i;
}

// ---- config / parameters -------------------------------------------------------------
void ref_config(double epsilon, int opt, int balanced, int with_idx, double rmax_scale, unsigned k) {
    config.epsilon = epsilon;
    config.opt = opt != 0;
    config.balanced = balanced != 0;
    config.with_rw_idx = with_idx != 0;
    config.rmax_scale = rmax_scale;
    config.k = k;
}
void ref_set_delta_pfail(double delta, double pfail) {
    config.delta = delta;
    config.pfail = pfail;
}
// which: 0 fora_setting, 1 fora_topk_setting, 2 montecarlo_setting, 3 bippr_setting, 4 fwdpush_setting
void ref_setting(int which, double* rmax, double* omega) {
    switch (which) {
        case 0: fora_setting(g_graph->n, g_graph->m); break;
        case 1: fora_topk_setting(g_graph->n, g_graph->m); break;
        case 2: montecarlo_setting(); break;
        case 3: bippr_setting(g_graph->n, g_graph->m); break;
        case 4: fwdpush_setting(g_graph->n, g_graph->m); break;
    }
    *rmax = config.rmax;
    *omega = config.omega;
}
void ref_get_params(double* out6) {
    out6[0] = config.alpha; out6[1] = config.epsilon; out6[2] = config.delta;
    out6[3] = config.pfail; out6[4] = config.rmax; out6[5] = config.omega;
}
void ref_set_rmax_omega(double rmax, double omega) { config.rmax = rmax; config.omega = omega; }

// ---- state init as done by query() (query.h:1427,1464-1467) / topk() (query.h:1343-1357)
void ref_init_query_state() {
    ppr.init_keys(g_graph->n);
    fwd_idx.first.nil = -1;
    fwd_idx.second.nil = -1;
    fwd_idx.first.initialize(g_graph->n);
    fwd_idx.second.initialize(g_graph->n);
    rw_counter.initialize(g_graph->n);
    bwd_idx.first.initialize(g_graph->n);
    bwd_idx.second.initialize(g_graph->n);
}
void ref_init_topk_state_fora() {
    int n = g_graph->n;
    fwd_idx.first.nil = -9; fwd_idx.first.initialize(n);
    fwd_idx.second.nil = -9; fwd_idx.second.initialize(n);
    rw_counter.nil = -9; rw_counter.init_keys(n);
    upper_bounds.nil = -9; upper_bounds.init_keys(n);
    lower_bounds.nil = -9; lower_bounds.init_keys(n);
    ppr.nil = -9; ppr.initialize(n);
    topk_filter.nil = -9; topk_filter.initialize(n);
}
// which: 2 montecarlo, 3 bippr, 4 fwdpush  (query.h:1361-1378)
void ref_init_topk_state_other(int which) {
    int n = g_graph->n;
    if (which == 2) { rw_counter.initialize(n); ppr.initialize(n); montecarlo_setting(); }
    else if (which == 3) { bippr_setting(n, g_graph->m); rw_counter.initialize(n);
                           bwd_idx.first.initialize(n); bwd_idx.second.initialize(n); ppr.initialize(n); }
    else if (which == 4) { fwdpush_setting(n, g_graph->m); fwd_idx.first.initialize(n);
                           fwd_idx.second.initialize(n); ppr.initialize(n); }
}

// ---- dumps -----------------------------------------------------------------------------
static void dump_imap(iMap<double>& m, double* dense, int* occur, int* n_occur) {
    int n = g_graph->n;
    if (dense) {
        for (int i = 0; i < n; ++i) dense[i] = 0.0;
        for (int i = 0; i < (int)m.occur.m_num; ++i) {
            int v = m.occur[i];
            if (m.exist(v)) dense[v] = m[v];
        }
    }
    if (occur) for (int i = 0; i < (int)m.occur.m_num; ++i) occur[i] = m.occur[i];
    if (n_occur) *n_occur = (int)m.occur.m_num;
}
void ref_dump_fwd(double* reserve, int* reserve_occur, int* n_reserve,
                  double* residue, int* residue_occur, int* n_residue) {
    dump_imap(fwd_idx.first, reserve, reserve_occur, n_reserve);
    dump_imap(fwd_idx.second, residue, residue_occur, n_residue);
}
void ref_dump_bwd(double* reserve, int* n_reserve, double* residue, int* n_residue) {
    dump_imap(bwd_idx.first, reserve, nullptr, n_reserve);
    dump_imap(bwd_idx.second, residue, nullptr, n_residue);
}
void ref_dump_ppr(double* dense) {
    int n = g_graph->n;
    for (int i = 0; i < n; ++i) dense[i] = 0.0;
    for (int i = 0; i < (int)ppr.occur.m_num; ++i) {
        int v = ppr.occur[i];
        if (ppr.exist(v)) dense[v] = ppr[v];
    }
}
void ref_counters(unsigned long long* total_rw, unsigned long long* hit_idx) {
    *total_rw = num_total_rw; *hit_idx = num_hit_idx;
}
void ref_reset_counters() { num_total_rw = 0; num_hit_idx = 0; num_iter_topk = 0; }
double ref_timer_used(int id) { return id < (int)Timer::timeUsed.size() ? Timer::used(id) : 0.0; }
void ref_timer_clear() { Timer::clearAll(); }

// ---- push ------------------------------------------------------------------------------
// forward_local_update_linear (algo.h:954-1018)
double ref_forward_push(int s, double rmax, double init_residual) {
    mute();
    double rsum = 1.0;
    forward_local_update_linear(s, *g_graph, rsum, rmax, init_residual);
    return rsum;
}
// forward_local_update_linear_topk (algo.h:1020-1093); state kept across calls.
static vector<int> g_forward_from;
static double g_rsum = 1.0;
void ref_push_topk_begin(int s) {
    mute();
    g_forward_from.clear();
    g_forward_from.push_back(s);
    fwd_idx.first.clean();
    fwd_idx.second.clean();
    g_rsum = 1.0;
    fwd_idx.second.insert(s, g_rsum);
}
double ref_push_topk_round(int s, double rmax, double lowest_rmax) {
    mute();
    forward_local_update_linear_topk(s, *g_graph, g_rsum, rmax, lowest_rmax, g_forward_from);
    return g_rsum;
}
int ref_push_topk_candidates(int* out) {
    if (out) for (size_t i = 0; i < g_forward_from.size(); ++i) out[i] = g_forward_from[i];
    return (int)g_forward_from.size();
}
// reverse_local_update_linear (algo.h:703-751); uses config.rmax
void ref_reverse_push(int t, double init_residual) {
    mute();
    reverse_local_update_linear(t, *g_graph, init_residual);
}

// ---- walks -----------------------------------------------------------------------------
void ref_random_walks(int start, long long count, int no_zero_hop, int* dest) {
    for (long long i = 0; i < count; ++i)
        dest[i] = no_zero_hop ? random_walk_no_zero_hop(start, *g_graph) : random_walk(start, *g_graph);
}

// ---- residue-seeded MC on a caller-provided push state -----------------------------------
// which: 0 compute_ppr_with_fwdidx (query.h:255), 1 _opt (query.h:334),
//        2 _topk (query.h:521), 3 _topk_with_bound (query.h:639), 4 compute_ppr_with_reserve
void ref_compute_ppr(int which, double rsum) {
    mute();
    switch (which) {
        case 0: compute_ppr_with_fwdidx(*g_graph, rsum); break;
        case 1: compute_ppr_with_fwdidx_opt(*g_graph, rsum); break;
        case 2: compute_ppr_with_fwdidx_topk(*g_graph, rsum); break;
        case 3: compute_ppr_with_fwdidx_topk_with_bound(*g_graph, rsum); break;
        case 4: compute_ppr_with_reserve(); break;
    }
}

// ---- whole-query drivers -----------------------------------------------------------------
// algo: 0 fora (fora_query_basic query.h:841), 2 montecarlo (query.h:16), 3 bippr (query.h:71),
//       4 fwdpush (query.h:1503-1508)
void ref_query(int algo, int s) {
    mute();
    if (algo == 0) fora_query_basic(s, *g_graph);
    else if (algo == 2) montecarlo_query(s, *g_graph);
    else if (algo == 3) bippr_query(s, *g_graph);
    else if (algo == 4) {
        Timer timer(FWD_LU);
        double rsum = 1;
        forward_local_update_linear(s, *g_graph, rsum, config.rmax);
        compute_ppr_with_reserve();
    }
}
// get_topk (query.h:1139) minus the precision bookkeeping; returns k pairs.
void ref_topk(int algo, int s, int* nodes, double* values) {
    mute();
    if (algo == 0) {
        if (config.opt) fora_query_topk_new(s, *g_graph);
        else fora_query_topk_with_bound(s, *g_graph);
    } else if (algo == 2) montecarlo_query_topk(s, *g_graph);
    else if (algo == 3) bippr_query_topk(s, *g_graph);
    else if (algo == 4) {
        double rsum = 1;
        forward_local_update_linear(s, *g_graph, rsum, config.rmax);
        compute_ppr_with_reserve();
    }
    topk_ppr();
    for (unsigned i = 0; i < config.k; ++i) { nodes[i] = topk_pprs[i].first; values[i] = topk_pprs[i].second; }
}
long ref_num_iter_topk() { return num_iter_topk; }

// exact top-k bookkeeping: install an exact list for source v, run compute_precision (algo.h:524)
void ref_set_exact_topk(int v, int count, const int* nodes, const double* values) {
    vector<pair<int, double> > lst(count);
    for (int i = 0; i < count; ++i) lst[i] = MP(nodes[i], values[i]);
    exact_topk_pprs[v] = lst;
}
void ref_set_topk_pprs(int count, const int* nodes, const double* values) {
    topk_pprs.resize(count);
    for (int i = 0; i < count; ++i) topk_pprs[i] = MP(nodes[i], values[i]);
}
void ref_compute_precision(int v, double* precision, double* recall) {
    mute();
    double p0 = result.topk_precision, r0 = result.topk_recall;
    compute_precision(v);
    *precision = result.topk_precision - p0;
    *recall = result.topk_recall - r0;
}

// ---- power iteration ground truth (query.h:1192-1224) -----------------------------------
void ref_power_iteration(int s, double* dense) {
    mute();
    unordered_map<int, double> map_ppr;
    fwd_power_iteration(*g_graph, s, map_ppr);
    for (int i = 0; i < g_graph->n; ++i) dense[i] = 0.0;
    for (auto& p : map_ppr) dense[p.first] = p.second;
}

// ---- index (build.h:302-366) --------------------------------------------------------------
// Runs the reference's build() with serialisation redirected to `folder` (must end with '/').
long long ref_build_index(const char* folder) {
    mute();
    config.graph_location = folder;
    build(*g_graph);
    return (long long)rw_idx.size();
}
void ref_load_index(const char* folder) {
    mute();
    config.graph_location = folder;
    deserialize_idx();
}
long long ref_index_size() { return (long long)rw_idx.size(); }
void ref_index_dump(unsigned long long* offsets, unsigned long long* counts, int* dest) {
    for (size_t v = 0; v < rw_idx_info.size(); ++v) { offsets[v] = rw_idx_info[v].first; counts[v] = rw_idx_info[v].second; }
    if (dest) for (size_t i = 0; i < rw_idx.size(); ++i) dest[i] = rw_idx[i];
}
void ref_index_set(int n, const unsigned long long* offsets, const unsigned long long* counts,
                   long long len, const int* dest) {
    rw_idx_info.resize(n);
    for (int v = 0; v < n; ++v) rw_idx_info[v] = MP(offsets[v], (unsigned long)counts[v]);
    rw_idx.assign(dest, dest + len);
}
const char* ref_index_file_names(int which) {
    static string s;
    s = which == 0 ? get_idx_file_name() : get_idx_info_name();
    return s.c_str();
}

// exact top-k archive (build.h:127-145)
void ref_save_exact_topk(const char* folder, const char* alias) {
    config.exact_pprs_folder = folder; config.graph_alias = alias;
    save_exact_topk_ppr();
}
int ref_load_exact_topk(const char* folder, const char* alias) {
    mute();
    config.exact_pprs_folder = folder; config.graph_alias = alias;
    exact_topk_pprs.clear();
    load_exact_topk_ppr();
    return (int)exact_topk_pprs.size();
}
int ref_get_exact_topk(int v, int cap, int* nodes, double* values) {
    auto it = exact_topk_pprs.find(v);
    if (it == exact_topk_pprs.end()) return -1;
    int c = min(cap, (int)it->second.size());
    for (int i = 0; i < c; ++i) { nodes[i] = it->second[i].first; values[i] = it->second[i].second; }
    return (int)it->second.size();
}

}  // extern "C"
