// fora_b200/csrc/bippr.cuh -- backward push and the BiPPR combine on sm_100a.
//
// Replaces reverse_local_update_linear (/root/reference/algo.h:703-751) and the per-target loop of
// bippr_query (/root/reference/query.h:90-113): for EVERY target t a backward push over the in-edges
//   reserve_t[v] += alpha*r;  every in-neighbour u of v:  residue_t[u] += ((1-alpha)*r)/d_out(u)
// followed by   ppr[t] = reserve_t[s] + sum_u count[u]/omega * residue_t[u].
// With the default epsilon the BiPPR r_max is 0.2-0.4, so one push touches the target and its
// in-neighbours only; the GPU runs n of them concurrently: a persistent grid, one CTA per target at
// a time, each CTA owning a private dense residue scratch that it cleans through its touched list
// (the reference's iMap idea, mylib.h:315-323).  Same frontier-synchronous schedule as the forward
// push: a vertex is pushed when its residue is >= r_max at level 0 (algo.h:725) and joins the next
// level when a scatter moves its residue from <= r_max to > r_max (algo.h:743).  The reference's
// early `break` (algo.h:725-726) and its double counting of re-inserted keys
// are deliberately not reproduced; the test suite checks both variants.
#pragma once
#include "common.cuh"

namespace fora {

constexpr int BWD_THREADS = 128;

template <typename OffT>
struct BwdArgs {
    int32_t n;
    double alpha, rmax, omega;
    int32_t source;                     // s of the query (reserve_t[s] is what BiPPR needs)
    const OffT* __restrict__ in_ptr;
    const int32_t* __restrict__ in_col;
    const int32_t* __restrict__ out_deg;
    const u64* __restrict__ counts;     // walk destination histogram of the query, may be null
    double* scratch_res;                // [nblocks*n] dense residue, all zero between targets
    double* scratch_rv;                 // [nblocks*n] residue snapshot per frontier entry
    int32_t* lists;                     // [nblocks*4*n] touched (n) | stamp (n) | cur (n) | nxt (n)
    int32_t* overflow;                  // [1] set if a touched list overflowed (cannot happen: one entry per vertex; kept as a guard)
    u32 epoch_base;                     // stamp value of target t_begin; target t uses epoch_base + (t - t_begin), never 0
    double* ppr;                        // [n] out, may be null
    int32_t t_begin, t_end;             // targets handled by this launch
    double* full_reserve;               // test hook: dense reserve of the single target, may be null
    int keep_residue;                   // test hook: leave the residue in scratch (block 0)
    u64* edges;                         // [1] counter
};

template <typename OffT>
__global__ void __launch_bounds__(BWD_THREADS) bippr_kernel(BwdArgs<OffT> a) {
    __shared__ int s_cnt[3];  // touched, cur, nxt sizes
    __shared__ double s_red[BWD_THREADS / WARP];
    __shared__ double s_acc;
    const size_t n = (size_t)a.n;
    double* res = a.scratch_res + n * blockIdx.x;
    double* rv = a.scratch_rv + n * blockIdx.x;
    // touched list: every vertex whose residue became non-zero for this target, listed ONCE -- stamp[v] holds the epoch of the
    // last target that listed v, so a frontier vertex that is zeroed in phase A and hit again in phase B is not re-appended
    // (it used to be: the list then grew with the sum of the frontier sizes over all levels and could overflow at small r_max).
    int32_t* touched = a.lists + 4 * n * blockIdx.x;
    u32* stamp = (u32*)(touched + n);
    int32_t* cur = touched + 2 * n;
    int32_t* nxt = cur + n;
    const int lane = lane_id(), w = threadIdx.x >> 5;
    u64 my_edges = 0;

    for (int32_t t = a.t_begin + blockIdx.x; t < a.t_end; t += gridDim.x) {
        const u32 ep = a.epoch_base + (u32)(t - a.t_begin);
        if (threadIdx.x == 0) {
            res[t] = 1.0; // init_residual (algo.h:718)
            touched[0] = t;
            stamp[t] = ep;
            cur[0] = t;
            s_cnt[0] = 1;
            s_cnt[1] = (1.0 < a.rmax) ? 0 : 1; // algo.h:725: the first popped vertex stops the loop if below rmax
            s_cnt[2] = 0;
            s_acc = 0.0;
        }
        __syncthreads();
        for (int level = 0; level < (1 << 20); ++level) {
            const int nc = s_cnt[1];
            if (nc == 0) break;
            // phase A: snapshot and zero
            for (int i = threadIdx.x; i < nc; i += BWD_THREADS) {
                const int32_t v = cur[i];
                const double r = res[v];
                res[v] = 0.0;
                rv[i] = r;
                if (a.full_reserve) a.full_reserve[v] += r * a.alpha;
                if (v == a.source) s_acc += r * a.alpha; // a vertex appears at most once per level: no race
            }
            __syncthreads();
            // phase B: warp per frontier vertex, lanes over its in-edges
            for (int i = w; i < nc; i += BWD_THREADS / WARP) {
                const int32_t v = cur[i];
                const double residual = (1 - a.alpha) * rv[i];
                const OffT b = a.in_ptr[v];
                const u32 d = (u32)(a.in_ptr[v + 1] - b);
                my_edges += (lane == 0) ? d : 0;
                for (u32 e0 = 0; e0 < d; e0 += WARP) {
                    const u32 e = e0 + lane;
                    bool first = false, cross = false;
                    int32_t u = 0;
                    if (e < d) {
                        u = a.in_col[b + (OffT)e];
                        const double inc = residual / (double)a.out_deg[u];
                        const double old = atomicAdd_block(&res[u], inc);
                        if (old == 0.0) first = atomicExch_block(&stamp[u], ep) != ep; // enters the touched list once per target
                        cross = !(old > a.rmax) && (old + inc > a.rmax);
                    }
                    const u32 mf = __ballot_sync(FULL, first), mc = __ballot_sync(FULL, cross);
                    if (mf) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd_block(&s_cnt[0], __popc(mf));
                        base = __shfl_sync(FULL, base, 0);
                        if (first) {
                            const size_t pos = (size_t)base + __popc(mf & lanemask_lt());
                            if (pos < n) touched[pos] = u;
                            else *a.overflow = 1;
                        }
                    }
                    if (mc) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd_block(&s_cnt[2], __popc(mc));
                        base = __shfl_sync(FULL, base, 0);
                        if (cross) nxt[base + __popc(mc & lanemask_lt())] = u;
                    }
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                s_cnt[1] = s_cnt[2];
                s_cnt[2] = 0;
            }
            int32_t* tmp = cur; cur = nxt; nxt = tmp;
            __syncthreads();
        }
        // combine and clean: every touched entry is read-and-zeroed exactly once
        const int nt = min(s_cnt[0], (int)n);
        double part = 0.0;
        if (!a.keep_residue) {
            for (int i = threadIdx.x; i < nt; i += BWD_THREADS) {
                const int32_t u = touched[i];
                const double r = __longlong_as_double((long long)atomicExch_block((u64*)&res[u], 0ull));
                if (a.counts && r != 0.0) part += (double)a.counts[u] * 1.0 / a.omega * r; // query.h:111
            }
        }
        part = warp_sum(part);
        if (lane == 0) s_red[w] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = s_acc;
            for (int i = 0; i < BWD_THREADS / WARP; ++i) tot += s_red[i];
            if (a.ppr) a.ppr[t] = tot;
        }
        __syncthreads();
        // cur/nxt may have been swapped an odd number of times: restore the canonical layout
        cur = touched + 2 * n;
        nxt = cur + n;
    }
    my_edges = warp_sum(my_edges);
    if (lane == 0 && my_edges && a.edges) atomicAdd(a.edges, my_edges);
}

} // namespace fora
