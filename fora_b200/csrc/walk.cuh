// fora_b200/csrc/walk.cuh -- residue-seeded Monte-Carlo phase and walk-index kernels on sm_100a.
//
// Replaces compute_ppr_with_fwdidx / _opt (/root/reference/query.h:255-413), the walk loops of
// random_walk / random_walk_no_zero_hop (algo.h:124-166), montecarlo_query (query.h:16-43) and
// the index build loop (build.h:344-354).
//
//   plan      per source v with residue r:  n_v = ceil(r/rsum*N), inc_v = (r/rsum*N/n_v)*rsum/N,
//             N = (ull)(omega*rsum)  -- the reference's expressions, evaluated in the same order in
//             IEEE double on the device (query.h:270,314-317; opt: 349,363-370), compacted in
//             ascending vertex order with an exclusive prefix sum of n_v (two deterministic passes)
//   walk      one walker per thread in a warp-converged loop (refill together, one Philox block = two steps together);
//             Philox4x32-10 keyed by (seed, query id, round) with counter (walk index within source, draw block,
//             source), so destinations do not depend on scheduling, chunk size, slot count or GPU count; row offsets
//             and neighbours through the read-only path; ppr[dest] += inc_v as fp64 RED.
#pragma once
#include "common.cuh"

namespace fora {

constexpr int PLAN_THREADS = 1024;
constexpr int PLAN_VPT = 4; // vertices per thread in the plan passes
constexpr int WALK_THREADS = 256;
#ifndef CFG_WALK_CHUNK
#define CFG_WALK_CHUNK 3840
#endif
// walks per chunk.  4 bytes of shared memory per walk; 8 resident CTAs x 3840 walks stay inside the 132 KB shared-memory
// carve-out (124 KB of L1 left).  Larger chunks = fewer chunk tails (lanes idle while the longest walks of a chunk
// finish): 1024: 98.6, 2048: 108.6, 2560: 110.9, 3840: 113.6, 4608 (next carve-out step): 102.3 G hops/s.
constexpr int WALK_CHUNK = CFG_WALK_CHUNK;
// A slot with few walks (a top-k round, a small graph) uses smaller chunks so that its walks still spread over the whole
// grid: chunk = walks / WALK_TARGET_CHUNKS rounded up to 256, within [256, WALK_CHUNK].  Evaluated identically by
// chunk_start_kernel and walk_kernel (and on every GPU of a split query) from the slot's walk count.
constexpr int WALK_TARGET_CHUNKS = 148 * 8 * 4;
__host__ __device__ __forceinline__ u32 walk_chunk_size(u64 walks) {
    const u64 c = ((walks / WALK_TARGET_CHUNKS) + 255) & ~255ull;
    return (u32)(c < 256 ? 256 : (c > (u64)WALK_CHUNK ? (u64)WALK_CHUNK : c));
}

struct PlanArgs {
    int32_t n;
    double alpha, omega;
    int opt;        // --opt: rsum *= (1-alpha), ppr[v] += alpha*r, r *= (1-alpha)  (query.h:349,363-364)
    int per_round;  // top-k rounds: 1: n_v = ceil(r*omega), inc = (r*omega/n_v)/omega (query.h:568-571); 3: the with-bound/index hybrid
    const double* __restrict__ residue; // [slots*n]
    double* ppr;                        // [slots*n] in/out: holds reserve on entry (may alias reserve)
    const double* __restrict__ rsum;    // [slots]
    const int32_t* __restrict__ slot_state;
    // outputs, per slot
    u32* __restrict__ blk_src;   // [slots*nblk]
    u64* __restrict__ blk_walk;  // [slots*nblk]
    int32_t* __restrict__ srcs;  // [slots*n]
    u64* __restrict__ woff;      // [slots*(n+1)]
    double* __restrict__ incs;   // [slots*n]
    u64* __restrict__ nsrc;      // [slots]
    u64* __restrict__ nwalk;     // [slots]
    int nblk;
    int no_credit; // multi-GPU walk split: parts > 0 start from a zero vector and skip the alpha*r credit
};

// n_v and inc_v for one source; identical expression order to query.h:314-317 (and 349,400-404).
__device__ __forceinline__ void plan_one(const PlanArgs& a, double r, double check_rsum, u64 num_random_walk, u64* n_v,
                                         double* inc) {
    if (a.per_round == 3) { // compute_ppr_with_fwdidx_topk_with_bound with index, query.h:659-662
        *n_v = (u64)ceil(__dmul_rn(r, a.omega));
        const double a_s = __ddiv_rn(__dmul_rn(__ddiv_rn(r, check_rsum), (double)num_random_walk), (double)*n_v);
        *inc = __ddiv_rn(__dmul_rn(a_s, check_rsum), (double)num_random_walk);
        return;
    }
    if (a.per_round) { // query.h:567-571
        const double num = ceil(__dmul_rn(r, a.omega));
        *n_v = (u64)num;
        const double a_s = __ddiv_rn(__dmul_rn(r, a.omega), (double)*n_v);
        *inc = __ddiv_rn(a_s, a.omega);
        return;
    }
    const double q = __dmul_rn(__ddiv_rn(r, check_rsum), (double)num_random_walk);
    *n_v = (u64)ceil(q);
    const double a_s = __ddiv_rn(q, (double)*n_v);
    *inc = __ddiv_rn(__dmul_rn(a_s, check_rsum), (double)num_random_walk);
}

// pass 1 (count) and pass 2 (fill) share the per-vertex evaluation; FILL selects the pass.  Every thread owns
// PLAN_VPT consecutive vertices (all their loads in flight together; one vertex per thread left the pass bound by
// block turnover: 151 K blocks of one load + two barriers each), so compaction keeps ascending vertex order.
template <bool FILL>
__global__ void __launch_bounds__(PLAN_THREADS) plan_kernel(PlanArgs a) {
    __shared__ u32 s_wsrc[PLAN_THREADS / WARP];
    __shared__ u64 s_wwalk[PLAN_THREADS / WARP];
    const int slot = blockIdx.y;
    if (a.slot_state[slot] == 0) return;
    const int v0 = (blockIdx.x * PLAN_THREADS + threadIdx.x) * PLAN_VPT;
    const size_t g0 = (size_t)slot * a.n;
    double check_rsum = a.rsum[slot];
    if (a.opt) check_rsum = __dmul_rn(check_rsum, 1.0 - a.alpha);
    const u64 num_random_walk = (u64)__dmul_rn(a.omega, check_rsum);
    const bool live = a.slot_state[slot] == 1;

    double r[PLAN_VPT], inc[PLAN_VPT];
    u64 n_v[PLAN_VPT];
#pragma unroll
    for (int k = 0; k < PLAN_VPT; ++k) r[k] = (live && v0 + k < a.n) ? a.residue[g0 + v0 + k] : 0.0;
    // fill pass with the alpha*r credit: the PPR words are fetched together with the residues, not behind the block scan's two barriers
    const bool credit_fill = FILL && (a.opt || a.per_round == 2) && !a.no_credit;
    double pold[PLAN_VPT];
#pragma unroll
    for (int k = 0; k < PLAN_VPT; ++k) pold[k] = (credit_fill && live && v0 + k < a.n) ? a.ppr[g0 + v0 + k] : 0.0;
    u32 flags = 0;
    u64 walks = 0;
#pragma unroll
    for (int k = 0; k < PLAN_VPT; ++k) {
        n_v[k] = 0;
        inc[k] = 0.0;
        if (r[k] > 0.0) {
            double rw = r[k];
            if (a.opt) rw = __dmul_rn(r[k], 1.0 - a.alpha);
            plan_one(a, rw, check_rsum, num_random_walk, &n_v[k], &inc[k]);
            ++flags;
            walks += n_v[k];
        }
    }
    // block-wide exclusive scan of (sources, walks) per thread: warp scan + scan of warp totals
    const int lane = lane_id(), w = threadIdx.x >> 5;
    const u32 fs = warp_incl_scan(flags);
    const u64 ws = warp_incl_scan64(walks);
    if (lane == 31) {
        s_wsrc[w] = fs;
        s_wwalk[w] = ws;
    }
    __syncthreads();
    if (w == 0) {
        u32 x = s_wsrc[lane];
        u64 y = s_wwalk[lane];
        const u32 xi = warp_incl_scan(x);
        const u64 yi = warp_incl_scan64(y);
        s_wsrc[lane] = xi - x;
        s_wwalk[lane] = yi - y;
        if (!FILL && lane == 31) {
            a.blk_src[(size_t)slot * a.nblk + blockIdx.x] = xi;
            a.blk_walk[(size_t)slot * a.nblk + blockIdx.x] = yi;
        }
    }
    if (!FILL) return;
    __syncthreads();
    if (flags) {
        size_t pos = (size_t)a.blk_src[(size_t)slot * a.nblk + blockIdx.x] + s_wsrc[w] + (fs - flags);
        u64 wo = a.blk_walk[(size_t)slot * a.nblk + blockIdx.x] + s_wwalk[w] + (ws - walks);
        const bool credit = (a.opt || a.per_round == 2) && !a.no_credit;
#pragma unroll
        for (int k = 0; k < PLAN_VPT; ++k) {
            if (!(r[k] > 0.0)) continue;
            a.srcs[g0 + pos] = v0 + k;
            a.woff[(size_t)slot * (a.n + 1) + pos] = wo;
            a.incs[g0 + pos] = inc[k];
            if (credit) a.ppr[g0 + v0 + k] = pold[k] + __dmul_rn(r[k], a.alpha); // query.h:363 / 562
            ++pos;
            wo += n_v[k];
        }
    }
}

// exclusive scan of the per-block totals of one slot (in place) + totals; one block per slot.
__global__ void __launch_bounds__(1024) plan_scan_kernel(PlanArgs a) {
    __shared__ u32 s_x[32];
    __shared__ u64 s_y[32];
    __shared__ u32 s_cx;
    __shared__ u64 s_cy;
    const int slot = blockIdx.x;
    if (a.slot_state[slot] == 0) return;
    u32* bs = a.blk_src + (size_t)slot * a.nblk;
    u64* bw = a.blk_walk + (size_t)slot * a.nblk;
    const int lane = lane_id(), w = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_cx = 0; s_cy = 0; }
    __syncthreads();
    for (int base = 0; base < a.nblk; base += 1024) {
        const int i = base + threadIdx.x;
        const u32 x = i < a.nblk ? bs[i] : 0;
        const u64 y = i < a.nblk ? bw[i] : 0;
        const u32 xi = warp_incl_scan(x);
        const u64 yi = warp_incl_scan64(y);
        if (lane == 31) { s_x[w] = xi; s_y[w] = yi; }
        __syncthreads();
        if (w == 0) {
            const u32 tx = s_x[lane];
            const u64 ty = s_y[lane];
            const u32 txi = warp_incl_scan(tx);
            const u64 tyi = warp_incl_scan64(ty);
            s_x[lane] = txi - tx;
            s_y[lane] = tyi - ty;
        }
        __syncthreads();
        const u32 cx = s_cx;
        const u64 cy = s_cy;
        if (i < a.nblk) {
            bs[i] = cx + s_x[w] + (xi - x);
            bw[i] = cy + s_y[w] + (yi - y);
        }
        __syncthreads();
        if (threadIdx.x == 1023) {
            s_cx = cx + s_x[w] + xi;
            s_cy = cy + s_y[w] + yi;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a.nsrc[slot] = s_cx;
        a.nwalk[slot] = s_cy;
        a.woff[(size_t)slot * (a.n + 1) + s_cx] = s_cy; // sentinel
    }
}

// chunk c of slot s starts inside source chunk_first[c]: largest i with woff[i] <= c * chunk size of the slot.
__global__ void __launch_bounds__(256) chunk_start_kernel(int32_t n, const u64* __restrict__ woff,
                                                          const u64* __restrict__ nsrc, const u64* __restrict__ nwalk,
                                                          u32* __restrict__ chunk_first, size_t chunk_cap,
                                                          const int32_t* __restrict__ slot_state) {
    const int slot = blockIdx.y;
    if (slot_state[slot] != 1) return;
    const u64 W = nwalk[slot];
    const u64 CH = walk_chunk_size(W);
    const u64 nchunks = (W + CH - 1) / CH;
    const u64* __restrict__ wo = woff + (size_t)slot * (n + 1);
    const u64 ns = nsrc[slot];
    for (u64 c = blockIdx.x * (u64)blockDim.x + threadIdx.x; c <= nchunks && c < chunk_cap; c += (u64)gridDim.x * blockDim.x) {
        const u64 target = c * CH;
        u64 lo = 0, hi = ns; // invariant: wo[lo] <= target; answer in [lo, hi)
        if (c == nchunks) {
            chunk_first[(size_t)slot * chunk_cap + c] = (u32)(ns ? ns - 1 : 0);
            continue;
        }
        while (hi - lo > 1) {
            const u64 mid = (lo + hi) >> 1;
            if (wo[mid] <= target) lo = mid;
            else hi = mid;
        }
        chunk_first[(size_t)slot * chunk_cap + c] = (u32)lo;
    }
}

struct WalkArgs {
    int32_t n;
    u32 alpha_thr;   // stop iff 32-bit draw < alpha * 2^32
    u32 seed_lo, seed_hi;
    int with_idx;
    const int32_t* __restrict__ srcs;
    const u64* __restrict__ woff;
    const double* __restrict__ incs;
    const u64* __restrict__ nsrc;
    const u64* __restrict__ nwalk;
    const u32* __restrict__ chunk_first;
    size_t chunk_cap;
    const int32_t* __restrict__ slot_state;
    const u32* __restrict__ qid;   // [slots] global query index (Philox key)
    u32 round_tag;
    double* ppr;                   // [slots*n]
    u64* __restrict__ hops;        // [slots]
    u64* __restrict__ idx_hits;    // [slots]
    // walk index (build.h): flat destinations + (offset,count) per vertex
    const u64* __restrict__ idx_off;
    const u64* __restrict__ idx_cnt;
    const int32_t* __restrict__ idx_dest;
    const u64* __restrict__ idx_used; // per-round cursor (top-k), may be null
    int all_idx;                      // shared walks (OUT_POOL instantiation): every walk is a hit in the pool, no walk is taken
    u32 part, nparts;                 // multi-GPU walk split: this launch walks chunks [part*C/nparts, (part+1)*C/nparts)
    int slot0;                        // first slot of this launch (blockIdx.y counts from it)
    u64 hot_elems;                    // HINT instantiation: neighbour slots from this position on are loaded with L2 evict_first
    int debug_no_red;                 // development ablation only (FORA_DEBUG_NO_RED): skip the ppr accumulation, results are WRONG
    // OUT != 0 (bulk walks: index build, Monte-Carlo, BiPPR, test hook): where the destinations go instead of ppr
    int32_t* out_dest;                // OUT_DEST: out_dest[global walk index] = destination (through new2old when set)
    u64* out_counts;                  // OUT_COUNT: out_counts[destination] += 1 (internal ids)
    const int32_t* __restrict__ new2old;
};
enum { OUT_PPR = 0, OUT_DEST = 1, OUT_COUNT = 2, OUT_POOL = 3 }; // OUT_POOL: shared walks, every walk of the launch is read from the wave's pool

// The walk itself.  Semantics of algo.h:124-166: a start with no out-edges returns itself; each
// step first stops with probability alpha (skipped once when NO_ZERO_HOP), then moves to a uniform
// out-neighbour, or back to THIS walk's start when the current vertex is dangling.
//
// Per CTA chunk of up to WALK_CHUNK walks (walk_chunk_size): (1) the prefix of the sources touching the chunk is staged in shared
// memory, (2) a divergence-free expansion pass resolves the owner source of every walk of the chunk
// (binary search, all lanes busy), (3) lanes fetch walks dynamically from a shared counter -- a lane that
// finishes a walk immediately starts the next one -- and advance two steps per Philox block (one
// Philox4x32-10 block = stop/pick, stop/pick).
//
// The loop is warp-converged: every iteration starts at an explicit reconvergence point, the lanes without a walk
// refill together, and then ALL lanes evaluate one Philox block and up to two steps together; a lane leaves the loop
// only when the whole warp has nothing left in this chunk.  (The first version let a lane that finished a walk go its
// own way -- fetch, next Philox block -- while its warp-mates were still stepping, and the warp never met again:
// 13.7 of 32 lanes active per issued instruction, 89 G hops/s; converged: 18.9 lanes, 97 G hops/s.)
// HINT: neighbour slots at positions >= hot_elems of the (hot-first) column array are loaded with L2 evict_first.
// OUT selects what happens with a destination: OUT_PPR ppr[dest] += inc_v (the query path), OUT_DEST store it at the walk's global
// index (index build, build.h:344-354; test hook), OUT_COUNT count it (montecarlo_query / bippr_query, query.h:25-31, 81-88).
template <typename OffT, bool NO_ZERO_HOP, bool HINT, int OUT = OUT_PPR>
__global__ void __launch_bounds__(WALK_THREADS) walk_kernel(WalkArgs a, CsrView<OffT> g) {
    // walk offsets of the sources touching the chunk, relative to the chunk start.  Entries 1.. lie in (0, chunk] and fit
    // 16 bits; entry 0 (<= 0, a source that started in an earlier chunk, possibly billions of walks ago) keeps 64 bits.
    // 4 bytes of shared memory per walk: the chunk can be large (few chunk tails) without eating the L1.
    static_assert(WALK_CHUNK + 1 <= 65535, "chunk offsets are staged as 16-bit values");
    __shared__ unsigned short s_rel[WALK_CHUNK + 2];
    __shared__ long long s_rel0;
    __shared__ unsigned short s_own[WALK_CHUNK];
    __shared__ u32 s_next;
    // OUT_COUNT: Monte-Carlo walks from ONE source end at a handful of vertices most of the time (alpha of them at the source itself),
    // and same-address global atomics serialise at ~1.5 ns each (9.4e8 walks per LJ-shape query: 0.28 s, measured).  A CTA counts the
    // destinations that claimed a line of this direct-mapped cache in shared memory and adds them to the histogram once at the end.
    constexpr int HC = OUT == OUT_COUNT ? 2048 : 1;
    __shared__ int32_t s_htag[HC];
    __shared__ u32 s_hcnt[HC];
    if (OUT == OUT_COUNT) {
        for (int i = threadIdx.x; i < HC; i += WALK_THREADS) { s_htag[i] = -1; s_hcnt[i] = 0; }
        __syncthreads();
    }
    const int slot = a.slot0 + (int)blockIdx.y;
    if (a.slot_state[slot] != 1) return;
    const u64 W = a.nwalk[slot];
    if (W == 0) return;
    const u32 CH = walk_chunk_size(W);
    const u64 nchunks = (W + CH - 1) / CH;
    const int32_t* __restrict__ srcs = a.srcs + (size_t)slot * a.n;
    const u64* __restrict__ woff = a.woff + (size_t)slot * (a.n + 1);
    const double* __restrict__ incs = a.incs + (size_t)slot * a.n;
    const u32* __restrict__ cfirst = a.chunk_first + (size_t)slot * a.chunk_cap;
    double* ppr = a.ppr + (size_t)slot * a.n;
    const u32 k0 = a.seed_lo ^ (a.qid[slot] * 0x9E3779B9u), k1 = a.seed_hi ^ a.round_tag;
    u64 my_hops = 0, my_hits = 0;
    u64 pol_stream = 0;
    if (HINT) pol_stream = l2_policy_evict_first();

    const u64 chunk_lo = a.nparts > 1 ? nchunks * a.part / a.nparts : 0;
    const u64 chunk_hi = a.nparts > 1 ? nchunks * (a.part + 1) / a.nparts : nchunks;
    for (u64 chunk = chunk_lo + blockIdx.x; chunk < chunk_hi; chunk += gridDim.x) {
        const u64 w0 = chunk * CH;
        const u32 nw = (u32)(min(W, w0 + (u64)CH) - w0);
        const u32 s_lo = cfirst[chunk], s_hi = cfirst[chunk + 1];
        const u32 cnt = s_hi - s_lo + 1; // sources touching this chunk, <= CH + 1
        __syncthreads();
        for (u32 i = threadIdx.x; i <= cnt; i += WALK_THREADS) {
            const long long rel = (long long)woff[s_lo + i] - (long long)w0;
            s_rel[i] = (unsigned short)max(0ll, min(rel, (long long)WALK_CHUNK + 1)); // entry 0 and the sentinel are never compared
            if (i == 0) s_rel0 = rel;
        }
        if (threadIdx.x == 0) s_next = 0;
        __syncthreads();
        for (u32 x = threadIdx.x; x < nw; x += WALK_THREADS) { // expansion: owner of walk x = last i with s_rel[i] <= x
            u32 lo = 0, hi = cnt;
            while (hi - lo > 1) {
                const u32 mid = (lo + hi) >> 1; // >= 1
                if ((u32)s_rel[mid] <= x) lo = mid;
                else hi = mid;
            }
            s_own[x] = (unsigned short)lo;
        }
        __syncthreads();

        if (OUT == OUT_POOL) {
            // shared walks: EVERY walk of the chunk is a hit in the wave's pool (idx_cnt[v] >= n_v by construction), so there is nothing
            // to walk and nothing to balance: walk x is read and added by thread x mod 256, four independent lookups in flight
            for (u32 x0 = threadIdx.x; x0 < nw; x0 += 4 * WALK_THREADS) {
                int32_t dest[4];
                double winc[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const u32 x = x0 + q * WALK_THREADS;
                    dest[q] = -1;
                    if (x < nw) {
                        const u32 own = s_own[x];
                        const u64 j = own ? (u64)(x - (u32)s_rel[own]) : (u64)((long long)x - s_rel0);
                        const int32_t v = srcs[s_lo + own];
                        winc[q] = incs[s_lo + own];
                        dest[q] = __ldcs(&a.idx_dest[a.idx_off[v] + j]);
                        ++my_hits;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (dest[q] >= 0) atomicAdd(&ppr[dest[q]], winc[q]);
            }
            continue;
        }

        int32_t cur = 0, start = 0;
        u32 jlo = 0, jhi = 0, blk = 0, myx = 0;
        double inc = 0.0;
        bool have = false, first = false, more = true;
        auto deliver = [&](int32_t dest, u32 x) { // one walk has ended at `dest`
            if (OUT == OUT_PPR) {
                if (!a.debug_no_red) atomicAdd(&ppr[dest], inc);
            } else if (OUT == OUT_DEST) {
                a.out_dest[w0 + x] = a.new2old ? a.new2old[dest] : dest;
            } else {
                const u32 hs = ((u32)dest * 0x9E3779B1u) >> 21; // 11 bits
                int32_t tag = s_htag[hs & (HC - 1)];
                if (tag == -1) tag = atomicCAS(&s_htag[hs & (HC - 1)], -1, dest) == -1 ? dest : s_htag[hs & (HC - 1)];
                if (tag == dest) atomicAdd(&s_hcnt[hs & (HC - 1)], 1u);
                else atomicAdd(&a.out_counts[dest], 1ull);
            }
        };
        for (;;) {
            __syncwarp();
            while (!have && more) { // refill; a walk resolved without walking (index hit / dangling start) fetches again
                const u32 x = atomicAdd(&s_next, 1u); // (one aggregated atomic per refilling warp measured 2.5 % slower)
                if (x >= nw) {
                    more = false;
                    break;
                }
                const u32 own = s_own[x];
                const u64 j = own ? (u64)(x - (u32)s_rel[own]) : (u64)((long long)x - s_rel0);
                const int32_t v = srcs[s_lo + own];
                if (OUT == OUT_PPR) inc = incs[s_lo + own];
                bool done = false;
                int32_t dest = v;
                if (a.with_idx) { // query.h:290-307: the first min(n_v, count) walks come from the index
                    const u64 used = a.idx_used ? a.idx_used[(size_t)slot * a.n + v] : 0;
                    const u64 avail = a.idx_cnt[v] - used;
                    if (j < avail) {
                        dest = a.idx_dest[a.idx_off[v] + used + j];
                        done = true;
                        ++my_hits;
                    }
                }
                if (!done && (u32)(g.ptr[v + 1] - g.ptr[v]) == 0) done = true; // algo.h:127-129
                if (done) {
                    deliver(dest, x);
                } else {
                    cur = start = v;
                    myx = x;
                    jlo = (u32)j;
                    jhi = (u32)(j >> 32);
                    blk = 0;
                    first = NO_ZERO_HOP;
                    have = true;
                }
            }
            if (!__any_sync(FULL, have)) break; // nobody holds a walk, so every lane has also seen the end of the chunk
            if (have) {
                // one Philox block = two steps; counter = (walk index lo, hi, block, source)
                const Philox4 rnd = philox4x32_10(jlo, jhi, blk++, (u32)start, k0, k1);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const u32 r_stop = half ? rnd.z : rnd.x;
                    const u32 r_pick = half ? rnd.w : rnd.y;
                    if (!first && r_stop < a.alpha_thr) { // algo.h:131-133
                        deliver(cur, myx);
                        have = false;
                        break;
                    }
                    first = false;
                    const OffT b = g.ptr[cur];
                    const u32 d = (u32)(g.ptr[cur + 1] - b);
                    if (d) {
                        const OffT pos = b + (OffT)__umulhi(r_pick, d); // algo.h:135-136
                        if (!HINT || (u64)pos < a.hot_elems) cur = __ldg(&g.col[pos]);
                        else cur = ld_s32_hint(&g.col[pos], pol_stream);
                        ++my_hops;
                    } else {
                        cur = start; // algo.h:138-140
                    }
                }
            }
        }
    }
    if (OUT == OUT_COUNT) {
        __syncthreads();
        for (int i = threadIdx.x; i < HC; i += WALK_THREADS)
            if (s_hcnt[i]) atomicAdd(&a.out_counts[s_htag[i]], (u64)s_hcnt[i]);
    }
    my_hops = warp_sum(my_hops);
    my_hits = warp_sum(my_hits);
    if (lane_id() == 0) {
        if (my_hops) atomicAdd(&a.hops[slot], my_hops);
        if (my_hits) atomicAdd(&a.idx_hits[slot], my_hits);
    }
}

// ppr[v] = counts[v] * 1.0 / omega   (query.h:35-38)
__global__ void counts_to_ppr_kernel(int32_t n, const u64* __restrict__ counts, double omega, double* __restrict__ ppr) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x)
        ppr[v] = counts[v] ? (double)counts[v] * 1.0 / omega : 0.0;
}

} // namespace fora
