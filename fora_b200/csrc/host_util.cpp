// fora_b200/csrc/host_util.cpp -- host-side parts of the C ABI (no GPU work): text loader,
// CSR construction, synthetic graph generator, parameter derivation.
//
// These replace Graph::init_nm / init_graph (/root/reference/graph.h:48-64,89-163) and the
// *_setting functions (/root/reference/algo.h:442-496).  Arithmetic follows the reference's
// expression order so the doubles are bit-identical (checked against the reference in tests/).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "fora_b200.h"

extern "C" {

int fora_host_read_attribute(const char* path, int32_t* n, int64_t* m) {
    // graph.h:48-64: skip to '=', read n; skip to '=', read m.
    FILE* f = fopen(path, "r");
    if (!f) return FORA_EIO;
    int c;
    long long mm = 0;
    int nn = 0;
    while ((c = fgetc(f)) != EOF && c != '=') {}
    if (fscanf(f, "%d", &nn) != 1) { fclose(f); return FORA_EIO; }
    while ((c = fgetc(f)) != EOF && c != '=') {}
    if (fscanf(f, "%lld", &mm) != 1) { fclose(f); return FORA_EIO; }
    fclose(f);
    *n = nn;
    *m = mm;
    return FORA_OK;
}

// Whitespace-separated decimal pairs, as fscanf("%d%d") reads them (graph.h:154): tokens pair up in sequence,
// line structure is irrelevant.  The reference's fscanf loop dominates load time at 1e9 edges, so the file is
// mapped once and scanned by all host cores: chunks are cut at whitespace, pass 1 counts the tokens of every chunk
// (which fixes the pairing parity at each chunk start), pass 2 parses the pairs of every chunk in parallel and the
// pieces are concatenated in file order (a pair straddling a chunk boundary is stitched in the merge).
namespace {
struct ChunkOut {
    std::vector<int32_t> src, dst;
    long long lead = 0, last = 0; // first token when the chunk starts mid-pair / dangling last token
    bool has_lead = false, has_last = false, bad = false;
};
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }
size_t count_tokens(const char* p, const char* e) {
    size_t c = 0;
    bool in = false;
    for (; p < e; ++p) {
        const bool d = is_digit(*p) || *p == '-';
        c += d && !in;
        in = d;
    }
    return c;
}
void parse_chunk(const char* p, const char* e, bool odd_start, int32_t n, bool store, ChunkOut& out) {
    long long val = 0, first = 0;
    bool in = false, neg = false, have_first = false;
    bool lead_pending = odd_start;
    auto token = [&](long long v) {
        if (lead_pending) { out.lead = v; out.has_lead = true; lead_pending = false; return; }
        if (!have_first) { first = v; have_first = true; return; }
        have_first = false;
        if (!(first < n) || !(v < n)) { out.bad = true; return; } // graph.h:155-156
        if (first == v) return;                                    // graph.h:157
        if (store) { out.src.push_back((int32_t)first); out.dst.push_back((int32_t)v); }
        else out.src.push_back(0); // count only
    };
    for (; p < e; ++p) {
        const char ch = *p;
        if (is_digit(ch)) { val = val * 10 + (ch - '0'); in = true; }
        else if (ch == '-' && !in) neg = true;
        else {
            if (in) token(neg ? -val : val);
            val = 0; in = false; neg = false;
        }
    }
    if (in) token(neg ? -val : val);
    if (have_first) { out.last = first; out.has_last = true; }
}
} // namespace

int64_t fora_host_read_edges(const char* path, int32_t n, int32_t* src, int32_t* dst) {
    FILE* f = fopen(path, "rb");
    if (!f) return FORA_EIO;
    fseek(f, 0, SEEK_END);
    const long long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)size + 1);
    if (size > 0 && fread(buf.data(), 1, (size_t)size, f) != (size_t)size) { fclose(f); return FORA_EIO; }
    fclose(f);
    buf[(size_t)size] = '\n';
    const char* base = buf.data();
    unsigned T = std::thread::hardware_concurrency();
    if (T == 0) T = 1;
    if ((long long)T * (1 << 16) > size) T = (unsigned)std::max<long long>(1, size >> 16);
    std::vector<size_t> cut(T + 1, (size_t)size);
    cut[0] = 0;
    for (unsigned t = 1; t < T; ++t) {
        size_t c = (size_t)(size * (long long)t / T);
        while (c < (size_t)size && (is_digit(base[c]) || base[c] == '-')) ++c; // never cut inside a token
        cut[t] = std::max(c, cut[t - 1]);
    }
    std::vector<size_t> ntok(T, 0);
    {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t) th.emplace_back([&, t] { ntok[t] = count_tokens(base + cut[t], base + cut[t + 1]); });
        for (auto& x : th) x.join();
    }
    std::vector<char> odd(T, 0);
    size_t acc = 0;
    for (unsigned t = 0; t < T; ++t) { odd[t] = (char)(acc & 1); acc += ntok[t]; }
    std::vector<ChunkOut> out(T);
    {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t)
            th.emplace_back([&, t] { parse_chunk(base + cut[t], base + cut[t + 1], odd[t] != 0, n, src != nullptr, out[t]); });
        for (auto& x : th) x.join();
    }
    int64_t kept = 0;
    bool bad = false;
    long long pend = 0;
    bool have_pend = false;
    for (unsigned t = 0; t < T; ++t) {
        bad = bad || out[t].bad;
        if (out[t].has_lead) { // second half of a pair that started in an earlier chunk
            if (have_pend) {
                if (!(pend < n) || !(out[t].lead < n)) bad = true;
                else if (pend != out[t].lead) {
                    if (src) { src[kept] = (int32_t)pend; dst[kept] = (int32_t)out[t].lead; }
                    ++kept;
                }
                have_pend = false;
            }
        }
        const size_t c = out[t].src.size();
        if (src && c) {
            memcpy(src + kept, out[t].src.data(), c * sizeof(int32_t));
            memcpy(dst + kept, out[t].dst.data(), c * sizeof(int32_t));
        }
        kept += (int64_t)c;
        if (out[t].has_last) { pend = out[t].last; have_pend = true; }
    }
    return bad ? (int64_t)FORA_ERANGE : kept;
}

// Stable counting sort == push_back in file order (graph.h:158-159).
int fora_host_csr_from_edges(int32_t n, int64_t ne, const int32_t* src, const int32_t* dst, int64_t* out_ptr,
                             int32_t* out_col, int64_t* in_ptr, int32_t* in_col) {
    if (n <= 0 || ne < 0) return FORA_EINVAL;
    std::fill(out_ptr, out_ptr + n + 1, 0);
    if (in_ptr) std::fill(in_ptr, in_ptr + n + 1, 0);
    for (int64_t e = 0; e < ne; ++e) {
        const int32_t s = src[e], d = dst[e];
        if ((uint32_t)s >= (uint32_t)n || (uint32_t)d >= (uint32_t)n) return FORA_ERANGE;
        if (s == d) continue;
        out_ptr[s + 1]++;
        if (in_ptr) in_ptr[d + 1]++;
    }
    for (int32_t i = 0; i < n; ++i) {
        out_ptr[i + 1] += out_ptr[i];
        if (in_ptr) in_ptr[i + 1] += in_ptr[i];
    }
    auto fill_dir = [&](const int32_t* key, const int32_t* val, const int64_t* ptr, int32_t* col) {
        std::vector<int64_t> pos(ptr, ptr + n);
        for (int64_t e = 0; e < ne; ++e) {
            if (src[e] == dst[e]) continue;
            col[pos[key[e]]++] = val[e];
        }
    };
    if (in_ptr && in_col) {
        std::thread t([&] { fill_dir(dst, src, in_ptr, in_col); });
        fill_dir(src, dst, out_ptr, out_col);
        t.join();
    } else {
        fill_dir(src, dst, out_ptr, out_col);
    }
    return FORA_OK;
}

// ---------------------------------------------------------------------------------------------
// Deterministic synthetic directed power-law graph: Chung-Lu style with a shifted
// power law weight (rank + shift)^(-1/(exponent-1)) on both endpoints (independent random
// relabelling of out- and in-ranks), `dangling_frac` of the vertices never emit an edge, self loops
// re-drawn so exactly m edges are kept, duplicates allowed.  Edge e is a pure function of
// (seed, e / CHUNK, e % CHUNK): the output does not depend on the thread count.
// ---------------------------------------------------------------------------------------------
namespace {
struct SplitMix {
    uint64_t s;
    explicit SplitMix(uint64_t seed) : s(seed) {}
    inline uint64_t next() {
        uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
    inline double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};
void permutation(std::vector<int32_t>& p, uint64_t seed) {
    SplitMix r(seed);
    for (size_t i = p.size(); i > 1; --i) {
        size_t j = (size_t)(r.next() % i);
        std::swap(p[i - 1], p[j]);
    }
}
} // namespace

int64_t fora_host_synth_edges(int32_t n, int64_t m, uint64_t seed, double exponent, double dangling_frac, int32_t* src,
                              int32_t* dst) {
    if (n < 4 || m < 1 || exponent <= 1.0) return FORA_EINVAL;
    const double beta = 1.0 / (exponent - 1.0);
    const double shift = 50.0;
    const double e1 = 1.0 - beta;
    // dangling set = the first n_d entries of a random permutation
    std::vector<int32_t> perm_out(n), perm_in(n);
    for (int32_t i = 0; i < n; ++i) perm_out[i] = perm_in[i] = i;
    permutation(perm_out, seed * 3 + 1);
    permutation(perm_in, seed * 3 + 2);
    int32_t n_d = (int32_t)(dangling_frac * n);
    if (n_d > n - 2) n_d = n - 2;
    const int32_t n_src = n - n_d; // out-ranks map onto perm_out[n_d ...]
    auto sample_rank = [&](double u, int32_t cnt) -> int32_t {
        const double a = pow(shift, e1), b = pow((double)cnt + shift, e1);
        double x = pow(u * (b - a) + a, 1.0 / e1) - shift;
        int32_t r = (int32_t)x;
        if (r < 0) r = 0;
        if (r >= cnt) r = cnt - 1;
        return r;
    };
    const int64_t CHUNK = 1 << 20;
    const int64_t nchunks = (m + CHUNK - 1) / CHUNK;
    unsigned T = std::thread::hardware_concurrency();
    if (T == 0) T = 1;
    if ((int64_t)T > nchunks) T = (unsigned)nchunks;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t) {
        th.emplace_back([&, t] {
            for (int64_t c = t; c < nchunks; c += T) {
                // chunk streams must not be shifted copies of each other: hash (seed, chunk) first
                SplitMix h(seed ^ (0xA0761D6478BD642FULL * (uint64_t)(c + 1)));
                h.next();
                SplitMix r(h.next() ^ (uint64_t)c);
                const int64_t lo = c * CHUNK, hi = std::min(m, lo + CHUNK);
                for (int64_t e = lo; e < hi; ++e) {
                    const int32_t s = perm_out[n_d + sample_rank(r.unit(), n_src)];
                    int32_t d;
                    do { d = perm_in[sample_rank(r.unit(), n)]; } while (d == s);
                    src[e] = s;
                    dst[e] = d;
                }
            }
        });
    }
    for (auto& x : th) x.join();
    return m;
}

// ---------------------------------------------------------------------------------------------
// algo.h:442-496
// ---------------------------------------------------------------------------------------------
int fora_host_setting(int which, int32_t n, int64_t m, double epsilon, double delta, double pfail, double alpha, int opt,
                      double rmax_scale, double* rmax_out, double* omega_out) {
    double rmax = 0.0, omega = 0.0;
    switch (which) {
        case 0: // fora_setting, algo.h:455-463
            rmax = epsilon * sqrt(delta / 3 / m / log(2 / pfail));
            if (opt) rmax *= rmax_scale / (1 - alpha);
            else rmax *= rmax_scale;
            omega = (2 + epsilon) * log(2 / pfail) / delta / epsilon / epsilon;
            break;
        case 1: // fora_topk_setting, algo.h:466-474
            rmax = epsilon * sqrt(delta / 3 / m / log(2 / pfail));
            rmax *= sqrt(1.0 * m * rmax) * rmax_scale * 3;
            omega = (2 + epsilon) * log(2 / pfail) / delta / epsilon / epsilon;
            break;
        case 2: // montecarlo_setting, algo.h:477-483
            omega = 3 * log(2 / pfail) / epsilon / epsilon / delta;
            break;
        case 3: // bippr_setting, algo.h:442-447
            rmax = epsilon * sqrt(m * 1.0 * delta / 3.0 / log(2.0 / pfail));
            rmax *= rmax_scale;
            omega = rmax * 3 * log(2.0 / pfail) / delta / epsilon / epsilon;
            break;
        case 4: // fwdpush_setting, algo.h:495
            rmax = rmax_scale * delta * epsilon * n / m;
            break;
        default:
            return FORA_EINVAL;
    }
    if (rmax_out) *rmax_out = rmax;
    if (omega_out) *omega_out = omega;
    return FORA_OK;
}

} // extern "C"
