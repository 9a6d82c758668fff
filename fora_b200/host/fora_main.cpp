// fora_b200/host/fora_main.cpp -- the `./fora` command line on top of the C ABI (include/fora_b200.h).
//
// Same surface as the reference's main() (/root/reference/fora.cpp:56-292): actions query / topk /
// batch-topk / build / gen-exact-topk / generate-ss-query, the flags --prefix --dataset --algo --epsilon
// --result_dir --exact_ppr_path --with_idx --rmax_scale --query_size --k --opt --balanced (plus the ones the
// reference parses and ignores), the same input files (attribute.txt, graph.txt, ssquery.txt), index and
// exact-top-k archives and result JSON.  All heavy work happens in libfora_b200.so on the GPU(s).
// Added, optional: --gpus N (shard queries / index sources over N GPUs), --seed S, --slots K.
#include <sys/resource.h>
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "archive.hpp"
#include "fora_b200.h"

using namespace std;
using namespace fora_host;

// ---------------------------------------------------------------------------------------------
// config (config.h:86-138 defaults)
// ---------------------------------------------------------------------------------------------
struct Config {
    string graph_alias = "nethept", graph_location, action, prefix = "d:\\dropbox\\research\\data\\", version = "vector";
    string exe_result_dir = "./", algo, exact_pprs_folder;
    bool multithread = false, with_rw_idx = false, opt = false, balanced = false, force_rebuild = false;
    double omega = 0, rmax = 0, pfail = 0, epsilon = 0, delta = 0, rmax_scale = 1, rw_cost_ratio = 8.0, alpha = 0.2;
    unsigned query_size = 1000, k = 500, hub_space_consum = 1;
    int gpus = 1, slots = 16;
    uint64_t seed = 0;
    bool shared_walks = false; // --shared_walks: the queries of a wave draw their walks from one pool (fora_ctx_set_shared_walks)
    bool split = false; // --split: the GPUs of --gpus answer every query TOGETHER (whole-graph SSPPR on huge graphs)
    string get_graph_folder() const { return prefix + graph_alias + "/"; }
};
static Config config;

struct GraphHost {
    int32_t n = 0;
    int64_t m = 0;
    vector<int32_t> src, dst; // edge list in file order (self loops already dropped); the CSR is built on the device
};

static bool exists_test(const string& name) {
    ifstream f(name.c_str());
    return f.good();
}
static string now_str() {
    time_t raw;
    time(&raw);
    char buf[80];
    strftime(buf, 80, "%Y-%m-%d %H:%M:%S", localtime(&raw));
    return buf;
}
static double wall() { return chrono::duration<double>(chrono::steady_clock::now().time_since_epoch()).count(); }

static void die(const string& msg, int code = 1) {
    cerr << msg << endl;
    exit(code);
}
#define CKF(ctx, call)                                                                     \
    do {                                                                                   \
        int rc__ = (call);                                                                 \
        if (rc__) die(string("fora_b200: ") + #call + ": " + fora_last_error(ctx), 1);     \
    } while (0)

// Graph graph(folder): graph.h:37-46,89-163 -- the text is parsed on all host cores, the adjacency (CSR) is built on the GPU
static GraphHost load_graph(const string& folder, bool need_edges) {
    GraphHost g;
    const string attr = folder + "attribute.txt";
    if (!exists_test(attr)) die("attribute file " + attr + " not find ");
    if (fora_host_read_attribute(attr.c_str(), &g.n, &g.m)) die("cannot parse " + attr);
    if (need_edges) {
        const string gf = folder + "graph.txt";
        if (!exists_test(gf)) die("graph file " + gf + " not find ");
        int64_t kept = fora_host_read_edges(gf.c_str(), g.n, nullptr, nullptr);
        if (kept == FORA_ERANGE) die("graph.txt: node id >= n (the reference asserts t1 < n, graph.h:155)");
        if (kept < 0) die("cannot read " + gf);
        g.src.resize((size_t)kept);
        g.dst.resize((size_t)kept);
        fora_host_read_edges(gf.c_str(), g.n, g.src.data(), g.dst.data());
    }
    cout << "init graph n: " << g.n << " m: " << g.m << endl;
    return g;
}

static vector<int> load_ss_query() { // algo.h:511-522
    const string filename = config.graph_location + "ssquery.txt";
    if (!exists_test(filename)) {
        cerr << "query file does not exist, please generate ss query files first" << endl;
        exit(0);
    }
    ifstream f(filename);
    vector<int> q;
    int v;
    while (f >> v) q.push_back(v);
    return q;
}

static void generate_ss_query(int n) { // algo.h:498-509
    const string filename = config.graph_location + "ssquery.txt";
    if (exists_test(filename)) {
        cout << "ss query set exists" << endl;
        return;
    }
    ofstream f(filename);
    for (unsigned i = 0; i < config.query_size; i++) f << rand() % n << endl;
}

// build.h:147-181
static string idx_name(const char* kind) {
    string f = config.graph_location + "randwalks.";
    if (config.rmax_scale != 1) f += to_string(config.rmax_scale) + ".";
    f += kind;
    if (config.opt) f += ".onehopopt";
    return f;
}
static string exact_topk_file() { // build.h:121-125
    if (config.exact_pprs_folder.empty() || config.exact_pprs_folder.back() != '/') config.exact_pprs_folder += "/";
    return config.exact_pprs_folder + config.graph_alias + ".topk.pprs";
}

// ---------------------------------------------------------------------------------------------
// one engine context per GPU
// ---------------------------------------------------------------------------------------------
struct Gpu {
    fora_ctx* ctx = nullptr;
};
static fora_group* g_group = nullptr; // --split: the contexts belong to one NCCL group
static vector<Gpu> open_gpus(const GraphHost& g, bool need_in) {
    vector<Gpu> gp((size_t)config.gpus);
    const uint64_t seed = config.seed ? config.seed : (uint64_t)time(nullptr); // the reference seeds from time(0)
    if (config.split && fora_group_create(config.gpus, nullptr, seed, &g_group)) die(string("fora_b200: ") + fora_group_last_error(nullptr));
    for (int d = 0; d < config.gpus; ++d) {
        if (g_group) gp[d].ctx = fora_group_ctx(g_group, d);
        else if (fora_ctx_create(d, seed, &gp[d].ctx)) die(string("fora_b200: ") + fora_last_error(nullptr));
        CKF(gp[d].ctx, fora_ctx_set_slots(gp[d].ctx, g_group ? 1 : config.slots));
        if (config.shared_walks) CKF(gp[d].ctx, fora_ctx_set_shared_walks(gp[d].ctx, 1));
        CKF(gp[d].ctx, fora_graph_build_from_edges(gp[d].ctx, g.n, g.m, g.src.data(), g.dst.data(), (int64_t)g.src.size(), need_in ? 1 : 0));
    }
    return gp;
}
static void close_gpus(vector<Gpu>& gp) {
    if (g_group) {
        fora_group_destroy(g_group);
        g_group = nullptr;
        return;
    }
    for (auto& x : gp) fora_ctx_destroy(x.ctx);
}
static void set_params_all(vector<Gpu>& gp) {
    fora_params p;
    memset(&p, 0, sizeof p);
    p.alpha = config.alpha; p.epsilon = config.epsilon; p.delta = config.delta; p.pfail = config.pfail;
    p.rmax = config.rmax; p.omega = config.omega; p.rmax_scale = config.rmax_scale;
    p.opt = config.opt; p.balanced = config.balanced; p.with_idx = config.with_rw_idx; p.k = config.k;
    for (auto& x : gp) CKF(x.ctx, fora_params_set(x.ctx, &p));
}
static void setting(int which, const GraphHost& g) { // algo.h:442-496
    fora_host_setting(which, g.n, g.m, config.epsilon, config.delta, config.pfail, config.alpha, config.opt, config.rmax_scale,
                      which == 2 ? nullptr : &config.rmax, which == 4 ? nullptr : &config.omega);
}
static int algo_id() {
    if (config.algo == "fora") return FORA_ALGO_FORA;
    if (config.algo == "fwdpush") return FORA_ALGO_FWDPUSH;
    if (config.algo == "montecarlo") return FORA_ALGO_MC;
    if (config.algo == "bippr") return FORA_ALGO_BIPPR;
    return -1;
}
static void load_index_all(vector<Gpu>& gp) { // deserialize_idx, build.h:194-207
    vector<int32_t> dest;
    vector<uint64_t> off, cnt;
    cout << "Index file name: " << idx_name("idx") << endl;
    try {
        load_index_dest(idx_name("idx"), dest);
        load_index_info(idx_name("info"), off, cnt);
    } catch (const exception& e) {
        die(e.what());
    }
    for (auto& x : gp) CKF(x.ctx, fora_index_upload(x.ctx, off.data(), cnt.data(), dest.data(), dest.size()));
}

// ---------------------------------------------------------------------------------------------
// result JSON (config.h:140-159,198-230,257-307): every leaf a quoted string like property_tree
// ---------------------------------------------------------------------------------------------
struct Result {
    double avg_query_time = 0, total_time = 0, rw_time = 0, push_time = 0, topk_time = 0, num_randwalk = 0, num_idx = 0;
    double precision = 0, recall = 0;
    int real_topk_source_count = 0;
};
static string jstr(double v) {
    ostringstream ss;
    ss.precision(17);
    ss << v;
    return ss.str();
}
static string jesc(const string& s) {
    string o;
    for (char c : s) {
        if (c == '"') o += "\\\"";
        else if (c == '\\') o += "\\\\";
        else if (c == '/') o += "\\/";
        else o += c;
    }
    return o;
}
static void save_json(const GraphHost& g, const Result& r, const string& start_time, int argc, char** argv) {
    string dir = config.exe_result_dir;
    if (dir.empty() || dir.back() != '/') dir += "/";
    dir += "execution/";
    for (size_t i = 1; i <= dir.size(); ++i)
        if (i == dir.size() || dir[i] == '/') mkdir(dir.substr(0, i).c_str(), 0777);
    string fn = dir + config.graph_alias + "." + config.action + "." + config.algo + "." + (config.with_rw_idx ? "with_idx" : "without_idx") + ".k-" +
                to_string(config.k) + ".rmax-" + to_string(config.rmax_scale) + ".json";
    string cmd;
    for (int i = 1; i < argc; ++i) cmd += string(" ") + argv[i];
    struct rusage ru;
    getrusage(RUSAGE_SELF, &ru);
    vector<pair<string, string> > cfg = {
        {"graph_alias", config.graph_alias}, {"action", config.action}, {"alpha", jstr(config.alpha)}, {"pfail", jstr(config.pfail)},
        {"epsilon", jstr(config.epsilon)}, {"delta", jstr(config.delta)}, {"idx", config.with_rw_idx ? "true" : "false"}, {"k", to_string(config.k)},
        {"rand-walk & push cost ratio", jstr(config.rw_cost_ratio)}, {"query-size", to_string(config.query_size)}, {"algo", config.algo},
        {"rmax", jstr(config.rmax)}, {"rmax-scale", jstr(config.rmax_scale)}, {"omega", jstr(config.omega)}, {"result-dir", dir}};
    const double cnt = r.real_topk_source_count;
    vector<pair<string, string> > res = {
        {"n", to_string(g.n)}, {"m", to_string(g.m)}, {"avg query time(s/q)", jstr(r.avg_query_time)},
        {"total memory usage(MB)", jstr(ru.ru_maxrss / 1000.0)}, {"total time usage(s)", jstr(r.total_time)},
        {"total time on rand-walks(s)", jstr(r.rw_time)}, {"total time on propagation(s)", jstr(r.push_time)},
        {"total time on sorting top-k ppr(s)", jstr(r.topk_time)},
        {"total time ratio on rand-walks(%)", jstr(r.total_time > 0 ? r.rw_time * 100 / r.total_time : 0)},
        {"total time ratio on propagation(%)", jstr(r.total_time > 0 ? r.push_time * 100 / r.total_time : 0)},
        {"total number of rand-walks", jstr(r.num_randwalk)}, {"total number of rand-walk idx used", jstr(r.num_idx)},
        {"total usage ratio of rand-walk idx", jstr(r.num_randwalk > 0 ? r.num_idx / r.num_randwalk : 0)},
        {"topk precision", jstr(cnt > 0 ? r.precision / cnt : NAN)}, {"topk recall", jstr(cnt > 0 ? r.recall / cnt : NAN)}};
    ofstream f(fn);
    f << "{\n    \"start_time\": \"" << start_time << "\",\n    \"end_time\": \"" << now_str() << "\",\n    \"command_line\": \"" << jesc(cmd) << "\",\n";
    auto dump = [&](const char* name, const vector<pair<string, string> >& kv, bool last) {
        f << "    \"" << name << "\": {\n";
        for (size_t i = 0; i < kv.size(); ++i)
            f << "        \"" << jesc(kv[i].first) << "\": \"" << jesc(kv[i].second) << "\"" << (i + 1 < kv.size() ? "," : "") << "\n";
        f << "    }" << (last ? "" : ",") << "\n";
    };
    dump("config", cfg, false);
    dump("result", res, false);
    // timer ids as in config.h:47-57: 0 top-k query, 3 FORA_QUERY, 5 FWD_LU, 6 RONDOM_WALK, 8 SORT_MAP
    vector<pair<string, string> > tm;
    if (config.action == "topk") tm.push_back({"0", jstr(r.total_time)});
    else tm.push_back({"3", jstr(r.total_time)});
    if (r.push_time > 0) tm.push_back({"5", jstr(r.push_time)});
    if (r.rw_time > 0) tm.push_back({"6", jstr(r.rw_time)});
    if (r.topk_time > 0) tm.push_back({"8", jstr(r.topk_time)});
    dump("timer", tm, true);
    f << "}\n";
}

// ---------------------------------------------------------------------------------------------
// sharded execution helpers
// ---------------------------------------------------------------------------------------------
struct Shard {
    int lo, hi;
};
static vector<Shard> shards(int count, int parts) {
    vector<Shard> s((size_t)parts);
    for (int p = 0; p < parts; ++p) {
        s[p].lo = (int)((int64_t)count * p / parts);
        s[p].hi = (int)((int64_t)count * (p + 1) / parts);
    }
    return s;
}

static void display_time_usage(const Result& r, unsigned query_size) { // algo.h:368-402
    cout << "Total cost (s): " << r.total_time << endl;
    if (config.algo != "fwdpush") cout << (r.total_time > 0 ? r.rw_time * 100.0 / r.total_time : 0) << "% for random walk cost" << endl;
    if (config.algo == "fora" || config.algo == "fwdpush") cout << (r.total_time > 0 ? r.push_time * 100.0 / r.total_time : 0) << "% for forward push cost" << endl;
    if (config.algo == "bippr") cout << (r.total_time > 0 ? r.push_time * 100.0 / r.total_time : 0) << "% for backward push cost" << endl;
    if (config.algo == "fora") cout << "-----------------------------" << endl;
    if (config.shared_walks && !config.with_rw_idx) cout << "Average shared rand-walk pool hit ratio: " << (r.num_randwalk > 0 ? r.num_idx * 100.0 / r.num_randwalk : 0) << "%" << endl;
    if (config.with_rw_idx) cout << "Average rand-walk idx hit ratio: " << (r.num_randwalk > 0 ? r.num_idx * 100.0 / r.num_randwalk : 0) << "%" << endl;
    if (config.action == "topk" && r.real_topk_source_count > 0) {
        cout << "Average top-K Precision: " << r.precision / r.real_topk_source_count << endl;
        cout << "Average top-K Recall: " << r.recall / r.real_topk_source_count << endl;
    }
    cout << "Average query time (s):" << r.total_time / query_size << endl;
    struct rusage ru;
    getrusage(RUSAGE_SELF, &ru);
    cout << "Memory usage (MB):" << ru.ru_maxrss / 1000.0 << endl << endl;
}

// ---------------------------------------------------------------------------------------------
// actions
// ---------------------------------------------------------------------------------------------
static Result do_query(const GraphHost& g, vector<Gpu>& gp) { // query(), query.h:1415-1515
    vector<int> queries = load_ss_query();
    unsigned query_size = min<unsigned>((unsigned)queries.size(), config.query_size);
    cout << "query_size=" << query_size << endl;
    const int which = config.algo == "fora" ? 0 : config.algo == "montecarlo" ? 2 : config.algo == "bippr" ? 3 : 4;
    setting(which, g);
    cout << "config.delta=" << config.delta << "\nconfig.pfail=" << config.pfail << "\nconfig.rmax=" << config.rmax << "\nconfig.omega=" << config.omega << endl;
    set_params_all(gp);
    for (unsigned i = 0; i < query_size; ++i)
        if (queries[i] < 0 || queries[i] >= g.n) die("query node out of range");
    vector<int32_t> src(queries.begin(), queries.begin() + query_size);
    vector<fora_query_stat> stats(query_size);
    vector<fora_batch_timing> tms(gp.size());
    if (g_group) { // --split: one query at a time, all GPUs on it (push on GPU 0, NCCL broadcast, walks split, NCCL all-reduce)
        if (config.algo != "fora") die("--split applies to --algo fora");
        Result r;
        const double t0s = wall();
        double bc = 0, rd = 0;
        for (unsigned i = 0; i < query_size; ++i) {
            fora_split_timing st;
            if (fora_group_query_split(g_group, src[i], i, nullptr, &stats[i], &st)) die(string("fora_b200: ") + fora_group_last_error(g_group));
            cout << i + 1 << ". source node:" << src[i] << endl;
            cout << "  split over " << st.n_gpus << " GPU(s): total " << st.total_ms << " ms = push " << st.push_ms << " + broadcast " << st.bcast_ms << " ("
                 << st.bcast_bytes << " B) + walks " << st.walk_ms << " + all-reduce " << st.reduce_ms << " (" << st.reduce_bytes << " B)" << endl;
            r.num_randwalk += (double)stats[i].n_walks;
            r.push_time += st.push_ms / 1e3;
            r.rw_time += st.walk_ms / 1e3;
            bc += st.bcast_ms / 1e3;
            rd += st.reduce_ms / 1e3;
        }
        r.total_time = wall() - t0s;
        r.avg_query_time = r.total_time / query_size;
        cout << "NCCL broadcast " << bc << " s, all-reduce " << rd << " s in total" << endl;
        config.query_size = query_size;
        display_time_usage(r, query_size);
        return r;
    }
    auto sh = shards((int)query_size, (int)gp.size());
    const double t0 = wall();
    vector<thread> th;
    for (size_t d = 0; d < gp.size(); ++d)
        th.emplace_back([&, d] {
            if (sh[d].hi > sh[d].lo) {
                CKF(gp[d].ctx, fora_ctx_set_query_base(gp[d].ctx, (uint64_t)sh[d].lo)); // Philox keyed by the global query index
                CKF(gp[d].ctx, fora_query_batch(gp[d].ctx, algo_id(), src.data() + sh[d].lo, sh[d].hi - sh[d].lo, nullptr, stats.data() + sh[d].lo, &tms[d]));
            }
        });
    for (auto& t : th) t.join();
    Result r;
    r.total_time = wall() - t0;
    for (unsigned i = 0; i < query_size; ++i) {
        cout << i + 1 << ". source node:" << src[i] << endl;
        r.num_randwalk += (double)stats[i].n_walks;
        r.num_idx += (double)stats[i].n_idx_hits;
    }
    for (auto& t : tms) { r.push_time = max(r.push_time, (double)t.push_ms / 1e3); r.rw_time = max(r.rw_time, (double)t.walk_ms / 1e3); }
    r.avg_query_time = r.total_time / query_size;
    config.query_size = query_size;
    display_time_usage(r, query_size);
    return r;
}

static void compute_precision(const vector<pair<int, double> >& est, const vector<pair<int, double> >& exact, unsigned k, double* precision, double* recall) {
    // algo.h:524-572: both ratios over |exact_map|
    map<int, double> topk_map, exact_map;
    for (auto& p : est) if (p.second > 0) topk_map.insert(p);
    const int size_e = (int)min<size_t>(k, exact.size());
    double rec = 0, pre = 0;
    for (int i = 0; i < size_e; ++i)
        if (exact[i].second > 0) {
            exact_map.insert(exact[i]);
            if (topk_map.count(exact[i].first)) rec++;
        }
    for (auto& p : topk_map) if (exact_map.count(p.first)) pre++;
    *recall = exact_map.empty() ? 0 : rec / exact_map.size();
    *precision = exact_map.empty() ? 0 : pre / exact_map.size();
}

static Result run_topk(const GraphHost& g, vector<Gpu>& gp, const vector<int32_t>& src, unsigned k, const ExactTopk& exact,
                       vector<vector<pair<int, double> > >* out_lists, double* avg_iters) {
    const unsigned nq = (unsigned)src.size();
    vector<int32_t> nodes((size_t)nq * k), iters(nq);
    vector<double> values((size_t)nq * k);
    vector<fora_query_stat> stats(nq);
    vector<fora_batch_timing> tms(gp.size());
    auto sh = shards((int)nq, (int)gp.size());
    const double t0 = wall();
    vector<thread> th;
    for (size_t d = 0; d < gp.size(); ++d)
        th.emplace_back([&, d] {
            if (sh[d].hi > sh[d].lo) {
                CKF(gp[d].ctx, fora_ctx_set_query_base(gp[d].ctx, (uint64_t)sh[d].lo)); // Philox keyed by the global query index
                CKF(gp[d].ctx, fora_topk_batch(gp[d].ctx, algo_id(), src.data() + sh[d].lo, sh[d].hi - sh[d].lo, k, nodes.data() + (size_t)sh[d].lo * k,
                                               values.data() + (size_t)sh[d].lo * k, iters.data() + sh[d].lo, stats.data() + sh[d].lo, &tms[d]));
            }
        });
    for (auto& t : th) t.join();
    Result r;
    r.total_time = wall() - t0;
    double it = 0;
    for (unsigned i = 0; i < nq; ++i) {
        vector<pair<int, double> > lst(k);
        for (unsigned j = 0; j < k; ++j) lst[j] = make_pair(nodes[(size_t)i * k + j], values[(size_t)i * k + j]);
        auto f = exact.find(src[i]);
        if (!exact.empty() && f != exact.end()) {
            double p, rc;
            compute_precision(lst, f->second, k, &p, &rc);
            r.precision += p; r.recall += rc; r.real_topk_source_count++;
        }
        if (out_lists) out_lists->push_back(lst);
        r.num_randwalk += (double)stats[i].n_walks;
        r.num_idx += (double)stats[i].n_idx_hits;
        it += iters[i];
    }
    for (auto& t : tms) r.topk_time = max(r.topk_time, (double)t.topk_ms / 1e3);
    if (avg_iters) *avg_iters = nq ? it / nq : 0;
    r.avg_query_time = nq ? r.total_time / nq : 0;
    return r;
}

static void prepare_topk_algo(const GraphHost& g, vector<Gpu>& gp) { // per-algo init of topk(), query.h:1343-1378
    if (config.algo == "fora") setting(0, g); // base parameters; the driver re-derives them per round (query.h:1002)
    else if (config.algo == "montecarlo") setting(2, g);
    else if (config.algo == "bippr") setting(3, g);
    else if (config.algo == "fwdpush") setting(4, g);
    set_params_all(gp);
}

static Result do_topk(const GraphHost& g, vector<Gpu>& gp, bool batch) { // topk() query.h:1309-1413 / batch_topk() 1517-1640
    vector<int> queries = load_ss_query();
    unsigned query_size = min<unsigned>((unsigned)queries.size(), config.query_size);
    if (!(config.k < (unsigned)g.n - 1 && config.k > 1)) die("Assertion `1 < k < n-1` failed");
    cout << "config.k=" << config.k << endl << "-----------------------------" << endl;
    ExactTopk exact;
    if (!load_exact_topk(exact_topk_file(), exact)) cout << "No exact topk ppr file " << exact_topk_file() << endl;
    prepare_topk_algo(g, gp);
    vector<int32_t> src(queries.begin(), queries.begin() + query_size);
    for (auto v : src) if (v < 0 || v >= g.n) die("query node out of range");
    vector<unsigned> ks;
    const unsigned step = config.k / 5;
    if (step > 0) for (unsigned i = 1; i < 5; ++i) ks.push_back(i * step);
    ks.push_back(config.k);
    Result r;
    double avg_it = 0;
    if (batch && config.algo == "fora") { // FORA re-runs the whole query set per k (query.h:1613-1636)
        map<unsigned, pair<double, double> > pr;
        const unsigned k_keep = config.k;
        for (unsigned k : ks) {
            config.k = k;
            set_params_all(gp);
            cout << "========================================\nk is set to be  config.k=" << k << endl;
            r = run_topk(g, gp, src, k, exact, nullptr, &avg_it);
            const double c = max(1, r.real_topk_source_count);
            pr[k] = make_pair(r.precision / c, r.recall / c);
            cout << "k=" << k << " precision=" << r.precision / c << " recall=" << r.recall / c << endl;
            cout << "Average query time (s):" << r.total_time / query_size << endl;
        }
        config.k = k_keep;
        cout << "-----------------------------\n" << config.algo << endl;
        for (unsigned k : ks) cout << k << "\t";
        cout << "\nPrecision:" << endl;
        for (unsigned k : ks) cout << pr[k].first << "\t";
        cout << "\nRecall:" << endl;
        for (unsigned k : ks) cout << pr[k].second << "\t";
        cout << endl;
        return r;
    }
    vector<vector<pair<int, double> > > lists;
    r = run_topk(g, gp, src, config.k, exact, &lists, &avg_it);
    for (unsigned i = 0; i < query_size; ++i) cout << i + 1 << ". source node:" << src[i] << endl << "-----------------------------" << endl;
    cout << "average iter times:" << (long)avg_it << endl;
    config.query_size = query_size;
    display_time_usage(r, query_size);
    if (config.algo != "fora") { // compute_precision_for_dif_k / display_precision_for_dif_k, algo.h:628-692
        cout << "-----------------------------\n" << config.algo << endl;
        for (unsigned k : ks) cout << k << "\t";
        vector<double> P, R;
        for (unsigned k : ks) {
            double ps = 0, rs = 0;
            int c = 0;
            for (unsigned i = 0; i < query_size; ++i) {
                auto f = exact.find(src[i]);
                if (f == exact.end()) continue;
                vector<pair<int, double> > pre(lists[i].begin(), lists[i].begin() + min<size_t>(k, lists[i].size()));
                double p, rc;
                compute_precision(pre, f->second, k, &p, &rc);
                ps += p; rs += rc; c++;
            }
            P.push_back(c ? ps / c : NAN);
            R.push_back(c ? rs / c : NAN);
        }
        cout << "\nPrecision:" << endl;
        for (double p : P) cout << p << "\t";
        cout << "\nRecall:" << endl;
        for (double p : R) cout << p << "\t";
        cout << endl;
    }
    return r;
}

static void do_build(const GraphHost& g, vector<Gpu>& gp) { // build(), build.h:302-366
    setting(0, g);
    set_params_all(gp);
    vector<uint64_t> off((size_t)g.n), cnt((size_t)g.n);
    uint64_t total = 0;
    CKF(gp[0].ctx, fora_index_info(gp[0].ctx, off.data(), cnt.data(), &total));
    cout << "tuned_index_size=" << total << endl;
    cout << "rand-walking..." << endl << "config.rmax=" << config.rmax << " config.omega=" << config.omega << " config.rmax*config.omega=" << config.rmax * config.omega << endl;
    vector<int32_t> dest((size_t)total);
    // shard by source range, balanced by the number of walks (not by node count)
    vector<int> cut(gp.size() + 1, g.n);
    cut[0] = 0;
    {
        size_t d = 1;
        for (int v = 0; v < g.n && d < gp.size(); ++v)
            if (off[v] >= total * d / gp.size()) cut[d++] = v;
    }
    const double t0 = wall();
    vector<thread> th;
    for (size_t d = 0; d < gp.size(); ++d)
        th.emplace_back([&, d] {
            if (cut[d + 1] > cut[d]) CKF(gp[d].ctx, fora_index_build(gp[d].ctx, off.data(), cnt.data(), cut[d], cut[d + 1], dest.data() + off[cut[d]]));
        });
    for (auto& t : th) t.join();
    cout << "index walks generated in " << wall() - t0 << " s on " << gp.size() << " GPU(s)" << endl;
    for (size_t d = 0; d < gp.size(); ++d) {
        uint64_t w = 0, hp = 0;
        double ms = 0;
        fora_index_build_stat(gp[d].ctx, &w, &hp, &ms);
        if (w) // index build moves 20 B per hop + 4 B per destination
            cout << "  GPU " << d << ": " << w << " walks, " << hp << " hops, walk kernels " << ms << " ms = " << hp / ms / 1e6 << " G hops/s, "
                 << (20.0 * hp + 4.0 * w) / ms / 1e6 << " GB/s algorithmic" << endl;
    }
    cout << "materializing..." << endl << "rw_idx.size()=" << dest.size() << " rw_idx_info.size()=" << off.size() << endl;
    try {
        save_index_dest(idx_name("idx"), dest);
        save_index_info(idx_name("info"), off, cnt);
    } catch (const exception& e) {
        die(e.what());
    }
    struct rusage ru;
    getrusage(RUSAGE_SELF, &ru);
    cout << "Memory usage (MB):" << ru.ru_maxrss / 1000.0 << endl << endl;
}

static void do_gen_exact_topk(const GraphHost& g, vector<Gpu>& gp) { // gen_exact_topk, query.h:1240-1307
    vector<int> queries = load_ss_query();
    unsigned query_size = min<unsigned>((unsigned)queries.size(), config.query_size);
    const string file = exact_topk_file();
    if (exists_test(file)) {
        cout << "exact top k exists" << endl;
        return;
    }
    if (!(config.k < (unsigned)g.n - 1 && config.k > 1)) die("Assertion `1 < k < n-1` failed");
    setting(0, g);
    set_params_all(gp);
    ExactTopk exact;
    vector<vector<pair<int, double> > > lists(query_size);
    auto sh = shards((int)query_size, (int)gp.size());
    const double t0 = wall();
    vector<thread> th;
    for (size_t d = 0; d < gp.size(); ++d)
        th.emplace_back([&, d] {
            vector<double> ppr((size_t)g.n), vals(config.k);
            vector<int32_t> nodes(config.k);
            for (int i = sh[d].lo; i < sh[d].hi; ++i) {
                CKF(gp[d].ctx, fora_power_iteration(gp[d].ctx, queries[i], 100, ppr.data())); // config.max_iter_num, config.h:115
                CKF(gp[d].ctx, fora_topk_of(gp[d].ctx, ppr.data(), config.k, nodes.data(), vals.data()));
                lists[i].resize(config.k);
                for (unsigned j = 0; j < config.k; ++j) lists[i][j] = make_pair(nodes[j], vals[j]);
            }
        });
    for (auto& t : th) t.join();
    cout << "average generation time (s): " << (wall() - t0) / max(1u, query_size) << endl;
    for (unsigned i = 0; i < query_size; ++i) exact[queries[i]] = lists[i];
    save_exact_topk(file, exact);
}

int main(int argc, char* argv[]) {
    ios::sync_with_stdio(false);
    const string start_time = now_str();
    cout << "\033[0;32m--------------start------------" << start_time << "\033[0m" << endl;
    {
        string c;
        for (int i = 1; i < argc; ++i) c += string(argv[i]) + " ";
        cout << "\033[0;32margs:" << c << "\033[0m" << endl;
    }
    srand((unsigned)time(nullptr));
    for (int i = 0; i < argc; i++)
        if (string(argv[i]) == "--help") {
            cout << "fora query --algo <algo> [options]\nfora topk  --algo <algo> [options]\nfora batch-topk --algo <algo> [options]\n"
                    "fora build [options]\nfora generate-ss-query [options]\nfora gen-exact-topk [options]\nfora\n\nalgo: \n  bippr\n  montecarlo\n  fora\n  fwdpush\n"
                    "options: \n  --prefix <prefix>\n  --epsilon <epsilon>\n  --dataset <dataset>\n  --query_size <queries count>\n  --k <top k>\n  --with_idx\n"
                    "  --exact_ppr_path <eaact-topk-pprs-path>\n  --rw_ratio <rand-walk cost ratio>\n  --result_dir <directory to place results>  --rmax_scale <scale of rmax>\n"
                    "  --opt\n  --balanced\n  --gpus <number of GPUs>\n  --split (all GPUs on one query at a time)\n  --shared_walks (the queries of a wave share one pool of random walks)\n  --seed <rng seed>\n  --slots <concurrent queries per GPU>\n"
                 << endl;
            exit(0);
        }
    if (argc < 2) die("sub command not regoznized");
    config.action = argv[1];
    cout << "action: " << config.action << endl;
    for (int i = 0; i < argc; i++) {
        const string arg = argv[i];
        auto val = [&](int j) -> const char* { return j < argc ? argv[j] : ""; };
        if (arg == "--prefix") config.prefix = val(i + 1);
        else if (arg == "--dataset") config.graph_alias = val(i + 1);
        else if (arg == "--algo") config.algo = val(i + 1);
        else if (arg == "--epsilon") config.epsilon = atof(val(i + 1));
        else if (arg == "--multithread") config.multithread = true;
        else if (arg == "--result_dir") config.exe_result_dir = val(i + 1);
        else if (arg == "--exact_ppr_path") config.exact_pprs_folder = val(i + 1);
        else if (arg == "--with_idx") config.with_rw_idx = true;
        else if (arg == "--rmax_scale") config.rmax_scale = atof(val(i + 1));
        else if (arg == "--force-rebuild") config.force_rebuild = true;
        else if (arg == "--query_size") config.query_size = (unsigned)atoi(val(i + 1));
        else if (arg == "--hub_space") config.hub_space_consum = (unsigned)atoi(val(i + 1));
        else if (arg == "--version") config.version = val(i + 1);
        else if (arg == "--k") config.k = (unsigned)atoi(val(i + 1));
        else if (arg == "--rw_ratio") config.rw_cost_ratio = atof(val(i + 1));
        else if (arg == "--opt") config.opt = true;
        else if (arg == "--balanced") config.balanced = true;
        else if (arg == "--gpus") config.gpus = max(1, atoi(val(i + 1)));
        else if (arg == "--split") config.split = true;
        else if (arg == "--shared_walks") config.shared_walks = true;
        else if (arg == "--seed") config.seed = strtoull(val(i + 1), nullptr, 10);
        else if (arg == "--slots") config.slots = max(1, atoi(val(i + 1)));
        else if (arg.substr(0, 2) == "--") {
            cerr << "command not recognize " << arg << endl;
            exit(1);
        }
    }
    const set<string> algos = {"bippr", "fora", "fwdpush", "montecarlo"};
    const bool needs_algo = config.action == "query" || config.action == "topk" || config.action == "batch-topk";
    if (needs_algo && !algos.count(config.algo)) {
        if (config.algo == "hubppr") die("hubppr is outside this build's scope (its index builder is not part of the reference either, README.md:107)");
        die("Wrong algo param:  config.algo=" + config.algo);
    }
    config.graph_location = config.get_graph_folder();
    Result result;
    if (config.action == "generate-ss-query") {
        GraphHost g = load_graph(config.graph_location, false);
        generate_ss_query(g.n);
    } else if (config.action == "query" || config.action == "topk" || config.action == "batch-topk" || config.action == "gen-exact-topk" || config.action == "build") {
        const bool need_in = config.algo == "bippr";
        GraphHost g = load_graph(config.graph_location, true);
        cout << "load graph finish" << endl;
        config.delta = 1.0 / g.n; // init_parameter, graph.h:173-183
        config.pfail = 1.0 / g.n;
        if (config.exact_pprs_folder.empty() || !exists_test(config.exact_pprs_folder)) config.exact_pprs_folder = config.graph_location;
        vector<Gpu> gp = open_gpus(g, need_in);
        if (config.with_rw_idx && config.action != "build" && config.action != "gen-exact-topk") {
            setting(0, g);
            load_index_all(gp);
        }
        if (config.action == "query") result = do_query(g, gp);
        else if (config.action == "topk") result = do_topk(g, gp, false);
        else if (config.action == "batch-topk") result = do_topk(g, gp, true);
        else if (config.action == "gen-exact-topk") do_gen_exact_topk(g, gp);
        else do_build(g, gp);
        close_gpus(gp);
        if (config.action == "query" || config.action == "topk") save_json(g, result, start_time, argc, argv);
        else cout << "\033[0;31m--------------stop------------" << now_str() << "\033[0m\n\n\n" << endl;
    } else {
        cerr << "sub command not regoznized" << endl;
        exit(1);
    }
    return 0;
}
