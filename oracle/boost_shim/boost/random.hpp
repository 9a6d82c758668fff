// Boost shim (oracle build only): random.hpp
// TEST INFRASTRUCTURE. Stand-in for the parts of Boost.Random the reference calls
// (/root/reference/algo.h:105-119): boost::taus88, boost::lagged_fibonacci607,
// boost::bernoulli_distribution<>, boost::variate_generator. Boost is not installed in this
// image; the generators are restated from their published definitions so that the
// reference's CPU cost per draw stays representative:
//   * taus88  - L'Ecuyer, "Maximally equidistributed combined Tausworthe generators",
//               Math. Comp. 65 (1996): three LFSRs (k,q,s) = (31,13,12), (29,2,4), (28,3,17),
//               xor-combined.
//   * lagged_fibonacci607 - additive lagged Fibonacci on [0,1) doubles with 48-bit mantissa,
//               lags (607, 273): x[i] = (x[i-607] + x[i-273]) mod 1.
// Stream-for-stream equality with a real Boost build is NOT claimed (and is moot: the
// reference seeds both from time(0)).
#ifndef SHIM_BOOST_RANDOM_HPP
#define SHIM_BOOST_RANDOM_HPP
#include <cstdint>

namespace boost {

namespace shim_detail {
template <int K, int Q, int S>
struct lfsr32 {
    uint32_t v;
    explicit lfsr32(uint32_t seed) {
        v = seed;
        if (v < (1u << (32 - K))) v += 1u << (32 - K);
    }
    inline uint32_t next() {
        const uint32_t b = ((v << Q) ^ v) >> (K - S);
        const uint32_t mask = 0xffffffffu << (32 - K);
        v = ((v & mask) << S) ^ b;
        return v;
    }
};
}  // namespace shim_detail

class taus88 {
    shim_detail::lfsr32<31, 13, 12> a;
    shim_detail::lfsr32<29, 2, 4> b;
    shim_detail::lfsr32<28, 3, 17> c;

public:
    typedef uint32_t result_type;
    explicit taus88(uint32_t seed = 341u) : a(seed), b(seed), c(seed) {}
    inline result_type operator()() { return a.next() ^ b.next() ^ c.next(); }
    static result_type min() { return 0; }
    static result_type max() { return 0xffffffffu; }
};

class lagged_fibonacci607 {
    enum { P = 607, Q = 273 };
    double x[P];
    unsigned i;

    void fill() {
        for (unsigned j = 0; j < Q; ++j) {
            double t = x[j] + x[j + (P - Q)];
            if (t >= 1.0) t -= 1.0;
            x[j] = t;
        }
        for (unsigned j = Q; j < P; ++j) {
            double t = x[j] + x[j - Q];
            if (t >= 1.0) t -= 1.0;
            x[j] = t;
        }
        i = 0;
    }

public:
    typedef double result_type;
    explicit lagged_fibonacci607(uint32_t seed = 331u) {
        // seed words from the minimal-standard LCG (16807 mod 2^31-1), two words per 48-bit value
        uint64_t s = seed % 2147483647u;
        if (s == 0) s = 1;
        const double two48 = 281474976710656.0;
        for (unsigned j = 0; j < P; ++j) {
            s = (s * 16807u) % 2147483647u;
            uint64_t lo = s;
            s = (s * 16807u) % 2147483647u;
            uint64_t hi = s;
            uint64_t bits = (lo | (hi << 31)) & 0xffffffffffffULL;
            x[j] = double(bits) / two48;
        }
        i = P;
    }
    inline result_type operator()() {
        if (i >= P) fill();
        return x[i++];
    }
    static result_type min() { return 0.0; }
    static result_type max() { return 1.0; }
};

template <class RealType = double>
class bernoulli_distribution {
    RealType p_;

public:
    typedef bool result_type;
    explicit bernoulli_distribution(const RealType& p = RealType(0.5)) : p_(p) {}
    template <class Engine>
    inline bool operator()(Engine& eng) const {
        if (p_ == RealType(0)) return false;
        return RealType(eng() - Engine::min()) <= p_ * RealType(Engine::max() - Engine::min());
    }
};

template <class EngineRef, class Dist>
class variate_generator;

template <class Engine, class Dist>
class variate_generator<Engine&, Dist> {
    Engine& e_;
    Dist d_;

public:
    variate_generator(Engine& e, Dist d) : e_(e), d_(d) {}
    inline typename Dist::result_type operator()() { return d_(e_); }
};

}  // namespace boost
#endif
