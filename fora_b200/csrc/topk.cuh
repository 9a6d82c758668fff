// fora_b200/csrc/topk.cuh -- top-k / k-th value selection and the dense power iteration.
//
// Replaces kth_ppr / topk_ppr (/root/reference/algo.h:578-610: nth_element / partial_sort_copy over
// the touched entries) with a radix select over the dense fp64 vectors of a whole wave of query slots: positive
// doubles order like their bit patterns, so the k-th largest value is found digit by digit.  Ties at the k-th value
// are broken towards the smaller node id by continuing the select over the id bits, which makes the result unique
// and reproducible (the reference leaves tie order unspecified).
// Slots beyond the number of positive entries stay (0, 0.0) as in algo.h:593-594.
//
// power_iteration_device restates fwd_power_iteration (query.h:1192-1224) densely.
#pragma once
#include "common.cuh"

namespace fora {

__device__ __forceinline__ u64 topk_key(double v) { return v > 0.0 ? (u64)__double_as_longlong(v) : 0ull; }

// ---------------------------------------------------------------------------------------------
// Batched select over MANY dense vectors (one per query slot): two passes over each vector and four launches for all of
// them (the first version selected slot by slot with 12 full passes and ~27 launches per vector, which left a top-k
// round launch- and sync-bound):
//   1. sel_hist_kernel      2048-bin histogram of the top 12 key bits (sign + exponent) of every positive entry
//   2. sel_pick_kernel      the bin holding the k-th largest entry, and how many entries lie above it
//   3. sel_classify_kernel  entries above the bin go straight to the output list, entries inside it to a candidate list
//   4. sel_finish_kernel    one CTA per vector finishes on the (small) candidate list: the remaining 52 key bits in
//                           8-bit digits, ties at the k-th value towards the smaller id, collect, optional bitonic sort
// ---------------------------------------------------------------------------------------------
constexpr int SEL_BINS = 2048;
struct SelSlot {
    u32 hist[SEL_BINS];
    u32 bin, above, n_positive, all_positive;
    u32 cand_count, out_count;
};
struct SelResult {
    u64 key_T;      // k-th largest key; 0 when the vector has fewer than k positive entries
    u32 out_count;  // entries in the output list (<= k)
    u32 all_positive;
};

__global__ void __launch_bounds__(256) sel_hist_kernel(const double* __restrict__ base, size_t stride, int32_t n,
                                                        const int32_t* __restrict__ slot_ids, SelSlot* st) {
    __shared__ u32 s_h[SEL_BINS];
    for (int b = threadIdx.x; b < SEL_BINS; b += blockDim.x) s_h[b] = 0;
    __syncthreads();
    const double* __restrict__ vals = base + stride * (size_t)slot_ids[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const u64 key = topk_key(vals[i]);
        if (key) atomicAdd(&s_h[key >> 52], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < SEL_BINS; b += blockDim.x)
        if (s_h[b]) atomicAdd(&st[blockIdx.y].hist[b], s_h[b]);
}

__global__ void __launch_bounds__(256) sel_pick_kernel(SelSlot* st, u32 k) {
    __shared__ u32 red[256];
    SelSlot& s = st[blockIdx.x];
    u32 t = 0;
    for (int b = threadIdx.x; b < SEL_BINS; b += blockDim.x) t += s.hist[b];
    red[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    s.n_positive = red[0];
    s.all_positive = red[0] < k;
    s.bin = 0;
    s.above = 0;
    if (!s.all_positive) {
        u32 rem = k;
        for (int b = SEL_BINS - 1; b >= 0; --b) {
            const u32 c = s.hist[b];
            if (c >= rem) { s.bin = (u32)b; s.above = k - rem; break; }
            rem -= c;
        }
    }
}

__global__ void __launch_bounds__(256) sel_classify_kernel(const double* __restrict__ base, size_t stride, int32_t n,
                                                            const int32_t* __restrict__ slot_ids, SelSlot* st,
                                                            u64* __restrict__ cand_keys, int32_t* __restrict__ cand_ids, size_t cand_stride,
                                                            int32_t* __restrict__ out_nodes, double* __restrict__ out_vals, u32 out_stride) {
    SelSlot& s = st[blockIdx.y];
    const double* __restrict__ vals = base + stride * (size_t)slot_ids[blockIdx.y];
    const bool all = s.all_positive;
    const u32 bin = s.bin;
    u64* ck = cand_keys + cand_stride * blockIdx.y;
    int32_t* ci = cand_ids + cand_stride * blockIdx.y;
    int32_t* on = out_nodes + (size_t)out_stride * blockIdx.y;
    double* ov = out_vals + (size_t)out_stride * blockIdx.y;
    const int iters = (n + WARP - 1) / WARP;
    for (int wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < iters; wi += (gridDim.x * blockDim.x) >> 5) {
        const int i = wi * WARP + lane_id();
        u64 key = 0;
        if (i < n) key = topk_key(vals[i]);
        const u32 top = (u32)(key >> 52);
        const bool win = key != 0 && (all || top > bin);
        const bool cand = key != 0 && !all && top == bin;
        const u32 mw = __ballot_sync(FULL, win), mc = __ballot_sync(FULL, cand);
        if (mw) {
            const int leader = __ffs(mw) - 1;
            u32 b0 = 0;
            if (lane_id() == leader) b0 = atomicAdd(&s.out_count, (u32)__popc(mw));
            b0 = __shfl_sync(FULL, b0, leader);
            const u32 pos = b0 + __popc(mw & lanemask_lt());
            if (win && pos < out_stride) { on[pos] = i; ov[pos] = __longlong_as_double((long long)key); }
        }
        if (mc) {
            const int leader = __ffs(mc) - 1;
            u32 b0 = 0;
            if (lane_id() == leader) b0 = atomicAdd(&s.cand_count, (u32)__popc(mc));
            b0 = __shfl_sync(FULL, b0, leader);
            const u32 pos = b0 + __popc(mc & lanemask_lt());
            if (cand) { ck[pos] = key; ci[pos] = i; }
        }
    }
}

// bitonic sort of `cnt` (value desc, id asc) pairs padded to a power of two `p2`, by one CTA
__device__ __forceinline__ void bitonic_sort_block(int32_t* nodes, double* vals, u32 cnt, u32 p2) {
    for (u32 i = cnt + threadIdx.x; i < p2; i += blockDim.x) { nodes[i] = 0x7fffffff; vals[i] = -1.0; }
    __syncthreads();
    for (u32 size = 2; size <= p2; size <<= 1) {
        for (u32 stride = size >> 1; stride > 0; stride >>= 1) {
            for (u32 t = threadIdx.x; t < (p2 >> 1); t += blockDim.x) {
                const u32 lo = 2 * t - (t & (stride - 1));
                const u32 hi = lo + stride;
                const bool desc_block = (lo & size) == 0;
                const double a = vals[lo], b = vals[hi];
                const int32_t ia = nodes[lo], ib = nodes[hi];
                const bool a_first = a > b || (a == b && ia < ib);
                if (a_first != desc_block) {
                    vals[lo] = b; vals[hi] = a;
                    nodes[lo] = ib; nodes[hi] = ia;
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(1024) sel_finish_kernel(SelSlot* st, SelResult* res, const u64* __restrict__ cand_keys,
                                                           const int32_t* __restrict__ cand_ids, size_t cand_stride,
                                                           int32_t* out_nodes, double* out_vals, u32 out_stride, u32 k, int sort) {
    __shared__ u32 h[256];
    __shared__ u64 sh_prefix;
    __shared__ u32 sh_rem, sh_eq, sh_idprefix, sh_out;
    SelSlot& s = st[blockIdx.x];
    const u64* __restrict__ ck = cand_keys + cand_stride * blockIdx.x;
    const int32_t* __restrict__ ci = cand_ids + cand_stride * blockIdx.x;
    int32_t* on = out_nodes + (size_t)out_stride * blockIdx.x;
    double* ov = out_vals + (size_t)out_stride * blockIdx.x;
    const u32 m = s.cand_count;
    u64 T = 0;
    if (!s.all_positive) {
        if (threadIdx.x == 0) { sh_prefix = (u64)s.bin << 52; sh_rem = k - s.above; sh_eq = 0; sh_out = s.out_count; }
        // remaining 52 key bits: six 8-bit digits (bits 51..4), then the last 4 bits
        for (int pass = 0; pass < 7; ++pass) {
            const int bits = pass < 6 ? 8 : 4, shift = pass < 6 ? 44 - 8 * pass : 0;
            const u64 hi_mask = ~0ull << (shift + bits);
            if (threadIdx.x < 256) h[threadIdx.x] = 0;
            __syncthreads();
            const u64 prefix = sh_prefix;
            for (u32 i = threadIdx.x; i < m; i += blockDim.x) {
                const u64 key = ck[i];
                if ((key & hi_mask) == prefix) atomicAdd(&h[(key >> shift) & ((1u << bits) - 1)], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                u32 rem = sh_rem;
                for (int b = (1 << bits) - 1; b >= 0; --b) {
                    const u32 c = h[b];
                    if (c >= rem) { sh_prefix |= (u64)b << shift; sh_eq = c; break; }
                    rem -= c;
                }
                sh_rem = rem; // after the last pass: how many entries equal to the k-th key belong to the top-k
            }
            __syncthreads();
        }
        T = sh_prefix;
        const u32 need_eq = sh_rem, count_eq = sh_eq;
        u32 idT = 0xffffffffu;
        if (need_eq != count_eq) { // ties at the k-th value: the need_eq smallest ids among them (ascending select over 32 bits)
            if (threadIdx.x == 0) { sh_idprefix = 0; sh_rem = need_eq; }
            for (int shift = 24; shift >= 0; shift -= 8) {
                const u32 hi_mask = shift >= 24 ? 0u : (~0u << (shift + 8));
                if (threadIdx.x < 256) h[threadIdx.x] = 0;
                __syncthreads();
                const u32 prefix = sh_idprefix;
                for (u32 i = threadIdx.x; i < m; i += blockDim.x)
                    if (ck[i] == T && ((u32)ci[i] & hi_mask) == prefix) atomicAdd(&h[((u32)ci[i] >> shift) & 0xff], 1u);
                __syncthreads();
                if (threadIdx.x == 0) {
                    u32 rem = sh_rem;
                    for (int b = 0; b < 256; ++b) {
                        const u32 c = h[b];
                        if (c >= rem) { sh_idprefix |= (u32)b << shift; break; }
                        rem -= c;
                    }
                    sh_rem = rem;
                }
                __syncthreads();
            }
            idT = sh_idprefix;
        }
        for (u32 i = threadIdx.x; i < m; i += blockDim.x) {
            const u64 key = ck[i];
            const int32_t id = ci[i];
            if (key > T || (key == T && (u32)id <= idT)) {
                const u32 pos = atomicAdd(&sh_out, 1u);
                if (pos < out_stride) { on[pos] = id; ov[pos] = __longlong_as_double((long long)key); }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s.out_count = sh_out;
    }
    __syncthreads();
    const u32 cnt = min(s.out_count, k);
    if (sort && cnt > 0) {
        u32 p2 = 1;
        while (p2 < cnt) p2 <<= 1;
        bitonic_sort_block(on, ov, cnt, p2);
    }
    if (threadIdx.x == 0) {
        res[blockIdx.x].key_T = s.all_positive ? 0ull : T;
        res[blockIdx.x].out_count = cnt;
        res[blockIdx.x].all_positive = s.all_positive;
    }
}

// ---------------------------------------------------------------------------------------------
// dense forward power iteration (query.h:1192-1224): iters synchronous sweeps,
// ppr[v] += alpha*r; out-neighbours += (1-alpha)*r/d_out; dangling mass -> the start node.
// buf = [ppr n][cur n][nxt n]
// ---------------------------------------------------------------------------------------------
template <typename OffT>
__global__ void __launch_bounds__(256) power_sweep_kernel(CsrView<OffT> g, int32_t n, int32_t start, double alpha,
                                                           double* __restrict__ ppr, const double* __restrict__ cur,
                                                           double* __restrict__ nxt) {
    const int lane = lane_id();
    for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < n; v += (gridDim.x * blockDim.x) >> 5) {
        const double r = cur[v];
        if (!(r > 0.0)) continue;
        if (lane == 0) ppr[v] += alpha * r;
        const OffT b = g.ptr[v];
        const u32 d = (u32)(g.ptr[v + 1] - b);
        const double remain = (1 - alpha) * r;
        if (d == 0) {
            if (lane == 0) atomicAdd(&nxt[start], remain);
        } else {
            const double avg = remain / (double)d;
            for (u32 e = lane; e < d; e += WARP) atomicAdd(&nxt[g.col[b + (OffT)e]], avg);
        }
    }
}

template <typename OffT>
static inline cudaError_t power_iteration_device(cudaStream_t stream, int num_sms, CsrView<OffT> g, int32_t n, int32_t start,
                                                 int iters, double alpha, double* buf, u64* launches) {
    double* ppr = buf;
    double* cur = buf + n;
    double* nxt = buf + 2 * (size_t)n;
    cudaError_t e;
    if ((e = cudaMemsetAsync(buf, 0, sizeof(double) * 3 * (size_t)n, stream)) != cudaSuccess) return e;
    const double one = 1.0;
    if ((e = cudaMemcpyAsync(cur + start, &one, sizeof(double), cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    for (int it = 0; it < iters; ++it) {
        power_sweep_kernel<OffT><<<num_sms * 8, 256, 0, stream>>>(g, n, start, alpha, ppr, cur, nxt);
        *launches += 1;
        std::swap(cur, nxt);
        if ((e = cudaMemsetAsync(nxt, 0, sizeof(double) * (size_t)n, stream)) != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

} // namespace fora

// =============================================================================================
// Bounds of the non --opt top-k driver: set_ppr_bounds (/root/reference/algo.h:1178-1261),
// calculate_lambda (algo.h:1169-1174) and the tail test of if_stop (algo.h:1147-1163).
// =============================================================================================
namespace fora {

__host__ __device__ inline double calculate_lambda(double rsum, double pfail, double upper_bound, double total_rw_num) {
    return 1.0 / 3 * log(2 / pfail) * rsum / total_rw_num +
           sqrt(4.0 / 9.0 * log(2.0 / pfail) * log(2.0 / pfail) * rsum * rsum + 8 * total_rw_num * log(2.0 / pfail) * rsum * upper_bound) / 2.0 /
               total_rw_num;
}

// one thread per vertex of one slot; bounds start at upper = 1, lower = 0 (query.h:940-941)
// set_ppr_bounds (algo.h:1178-1261) for every listed slot in ONE launch: blockIdx.y = position in `slots`; the per-slot scalars
// (rsum, number of walks of the round) are read from the slot block on the device, so no host round trip precedes the launch.
// A slot takes part when delta < threshold, it walked this round and has residue left (query.h:745-746).
__global__ void __launch_bounds__(256) ppr_bounds_kernel(int32_t n, size_t stride, const int32_t* __restrict__ slots, const double* __restrict__ rsum_of,
                                                          const u64* __restrict__ nwalk_of, double pfail, double delta, double threshold,
                                                          const double* __restrict__ ppr_base, const double* __restrict__ reserve_base,
                                                          double* __restrict__ upper_base, double* __restrict__ lower_base) {
    const int slot = slots[blockIdx.y];
    const double rsum = rsum_of[slot], total_rw_num = (double)nwalk_of[slot];
    if (!(delta < threshold && total_rw_num > 0 && rsum > 0)) return;
    const double* __restrict__ ppr = ppr_base + stride * slot;
    const double* __restrict__ reserve = reserve_base + stride * slot;
    double* __restrict__ upper = upper_base + stride * slot;
    double* __restrict__ lower = lower_base + stride * slot;
    const double min_ppr = 1.0 / n, sqrt_min_ppr = sqrt(1.0 / n);
    const double epsilon_v_div = sqrt(2.67 * rsum * log(2.0 / pfail) / total_rw_num);
    const double default_epsilon_v = epsilon_v_div / sqrt_min_ppr;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
        const double p = ppr[v];
        if (p <= 0) continue;
        double res = reserve[v]; // 0 when absent
        double epsilon_a;
        if (upper[v] > res) epsilon_a = calculate_lambda(rsum, pfail, upper[v] - res, total_rw_num);
        else epsilon_a = calculate_lambda(rsum, pfail, 1 - res, total_rw_num);
        const double ub_eps_a = p + epsilon_a;
        double lb_eps_a = p - epsilon_a;
        if (!(lb_eps_a > 0)) lb_eps_a = 0;
        double epsilon_v = default_epsilon_v;
        if (res > min_ppr) {
            res = fmax(res, lower[v]);
            epsilon_v = epsilon_v_div / sqrt(res);
        } else if (lower[v] > 0) {
            epsilon_v = epsilon_v_div / sqrt(lower[v]);
        }
        double ub_eps_v = 1.0, lb_eps_v = 0.0;
        if (1.0 - epsilon_v > 0) {
            ub_eps_v = p / (1.0 - epsilon_v);
            lb_eps_v = p / (1.0 + epsilon_v);
        }
        const double up_bound = fmin(fmin(ub_eps_a, ub_eps_v), 1.0);
        const double low_bound = fmax(fmax(lb_eps_a, lb_eps_v), res);
        if (up_bound > 0) upper[v] = up_bound;
        if (low_bound >= 0) lower[v] = low_bound;
    }
}

__global__ void fill_kernel(double* p, size_t n, double v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// if_stop part 2 (algo.h:1124-1163) for every listed slot in one launch each (blockIdx.y = position j in the list; vector j of the
// preceding batched select holds the k nodes with the largest lower bounds of slot slots[j]):
//   fail[2j]   |= any of them has upper/lower > 1+eps;   in_topk (one byte map per slot) marks them
__global__ void stop_mark_kernel(const int32_t* __restrict__ nodes_base, u32 node_stride, u32 k, const int32_t* __restrict__ slots, size_t stride,
                                 const double* __restrict__ upper_base, const double* __restrict__ lower_base, double eps,
                                 unsigned char* __restrict__ in_topk_base, u32* fail) {
    const int j = blockIdx.y, slot = slots[j];
    const int32_t* __restrict__ nodes = nodes_base + (size_t)node_stride * j;
    const double* __restrict__ upper = upper_base + stride * slot;
    const double* __restrict__ lower = lower_base + stride * slot;
    unsigned char* __restrict__ in_topk = in_topk_base + stride * slot;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
        const int32_t v = nodes[i];
        in_topk[v] = 1;
        if (upper[v] / lower[v] > 1.0 + eps) atomicOr(fail + 2 * j, 1u);
    }
}
//   fail[2j+1] |= a node outside the top-k with ppr > 0 whose upper bound exceeds low_bound_k*(1+eps) without being
//   separated by (1+eps)/(1-eps)
__global__ void __launch_bounds__(256) stop_tail_kernel(int32_t n, const int32_t* __restrict__ slots, size_t stride, const double* __restrict__ ppr_base,
                                                         const double* __restrict__ upper_base, const double* __restrict__ lower_base,
                                                         const unsigned char* __restrict__ in_topk_base, const double* __restrict__ low_bound_k_of,
                                                         double eps, u32* fail) {
    const int j = blockIdx.y, slot = slots[j];
    const double* __restrict__ ppr = ppr_base + stride * slot;
    const double* __restrict__ upper = upper_base + stride * slot;
    const double* __restrict__ lower = lower_base + stride * slot;
    const unsigned char* __restrict__ in_topk = in_topk_base + stride * slot;
    const double low_bound_k = low_bound_k_of[j];
    bool bad = false;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
        if (in_topk[v] || ppr[v] <= 0) continue;
        const double u = upper[v], l = lower[v];
        if (u > low_bound_k * (1.0 + eps) && !(u > (1 + eps) / (1 - eps) * l)) bad = true;
    }
    if (__any_sync(FULL, bad) && lane_id() == 0) atomicOr(fail + 2 * j + 1, 1u);
}
__global__ void stop_unmark_kernel(const int32_t* __restrict__ nodes_base, u32 node_stride, u32 k, const int32_t* __restrict__ slots, size_t stride,
                                   unsigned char* __restrict__ in_topk_base) {
    const int j = blockIdx.y;
    const int32_t* __restrict__ nodes = nodes_base + (size_t)node_stride * j;
    unsigned char* __restrict__ in_topk = in_topk_base + stride * slots[j];
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) in_topk[nodes[i]] = 0;
}

} // namespace fora
