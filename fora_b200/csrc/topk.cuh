// fora_b200/csrc/topk.cuh -- top-k / k-th value selection and the dense power iteration.
//
// Replaces kth_ppr / topk_ppr (/root/reference/algo.h:578-610: nth_element / partial_sort_copy over
// the touched entries) with a radix select over the dense fp64 vector: positive doubles order like
// their bit patterns, so the k-th largest value is found digit by digit (8 passes of 8 bits, one
// 256-bin shared-memory histogram per block per pass).  Ties at the k-th value are broken towards
// the smaller node id by continuing the select over the id bits, which makes the result unique
// and reproducible (the reference leaves tie order unspecified).
// Slots beyond the number of positive entries stay (0, 0.0) as in algo.h:593-594.
//
// power_iteration_device restates fwd_power_iteration (query.h:1192-1224) densely.
#pragma once
#include "common.cuh"

namespace fora {

struct TopkWork {};

struct SelectState {
    u64 prefix;     // high bits of the key decided so far
    u64 remaining;  // how many elements still to take inside the current prefix
    u64 n_positive;
    u64 key_T;      // k-th largest key (valid after the value passes)
    u64 need_eq;    // how many elements with key == key_T belong to the top-k
    u64 count_eq;
    u32 id_prefix;  // tie-break select over ids (ascending)
    u32 id_T;       // largest id taken among the ties
    u32 hist[256];
    u32 out_count;
    u32 all_positive; // fewer than k positive entries: take them all
};

__device__ __forceinline__ u64 topk_key(double v) { return v > 0.0 ? (u64)__double_as_longlong(v) : 0ull; }

// histogram of digit `shift` among elements whose key matches `prefix` on the bits above it
__global__ void __launch_bounds__(256) topk_hist_kernel(const double* __restrict__ vals, int32_t n, SelectState* st, int shift) {
    __shared__ u32 s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const u64 prefix = st->prefix;
    const u64 hi_mask = shift >= 56 ? 0ull : (~0ull << (shift + 8));
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const u64 key = topk_key(vals[i]);
        if (key != 0 && (key & hi_mask) == prefix) atomicAdd(&s_h[(key >> shift) & 0xff], 1u);
    }
    __syncthreads();
    if (s_h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], s_h[threadIdx.x]);
}

// pick the digit that contains the `remaining`-th largest element, descend into it
__global__ void topk_pick_kernel(SelectState* st, int shift, u64 k) {
    if (threadIdx.x != 0) return;
    if (shift == 56) { // first pass: initialise
        u64 total = 0;
        for (int b = 0; b < 256; ++b) total += st->hist[b];
        st->n_positive = total;
        st->remaining = k;
        st->all_positive = total < k;
    }
    if (!st->all_positive) {
        u64 rem = st->remaining;
        for (int b = 255; b >= 0; --b) {
            const u64 c = st->hist[b];
            if (c >= rem) {
                st->prefix |= (u64)b << shift;
                if (shift == 0) { st->count_eq = c; st->need_eq = rem; st->key_T = st->prefix; }
                break;
            }
            rem -= c;
        }
        st->remaining = rem;
    } else if (shift == 0) {
        st->key_T = 1; // every positive key is > 0: take all of them
        st->need_eq = 0;
        st->count_eq = 0;
    }
    for (int b = 0; b < 256; ++b) st->hist[b] = 0;
}

// tie-break: among key == key_T select the need_eq smallest ids (ascending radix select over 32 bits)
__global__ void __launch_bounds__(256) topk_idhist_kernel(const double* __restrict__ vals, int32_t n, SelectState* st, int shift) {
    __shared__ u32 s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    if (st->all_positive || st->need_eq == st->count_eq) return;
    const u64 T = st->key_T;
    const u32 prefix = st->id_prefix;
    const u32 hi_mask = shift >= 24 ? 0u : (~0u << (shift + 8));
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (topk_key(vals[i]) == T && ((u32)i & hi_mask) == prefix) atomicAdd(&s_h[((u32)i >> shift) & 0xff], 1u);
    }
    __syncthreads();
    if (s_h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], s_h[threadIdx.x]);
}
__global__ void topk_idpick_kernel(SelectState* st, int shift) {
    if (threadIdx.x != 0) return;
    if (st->all_positive || st->need_eq == st->count_eq) {
        st->id_T = 0xffffffffu;
        return;
    }
    if (shift == 24) st->remaining = st->need_eq;
    u64 rem = st->remaining;
    for (int b = 0; b < 256; ++b) {
        const u64 c = st->hist[b];
        if (c >= rem) {
            st->id_prefix |= (u32)b << shift;
            break;
        }
        rem -= c;
    }
    st->remaining = rem;
    if (shift == 0) st->id_T = st->id_prefix;
    for (int b = 0; b < 256; ++b) st->hist[b] = 0;
}

__global__ void __launch_bounds__(256) topk_collect_kernel(const double* __restrict__ vals, int32_t n, SelectState* st,
                                                            int32_t* __restrict__ out_nodes, double* __restrict__ out_vals, u32 cap) {
    const u64 T = st->key_T;
    const u32 idT = st->id_T;
    const bool all = st->all_positive;
    const int iters = (n + WARP - 1) / WARP;
    for (int wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < iters; wi += (gridDim.x * blockDim.x) >> 5) {
        const int i = wi * WARP + lane_id();
        bool take = false;
        double v = 0.0;
        if (i < n) {
            v = vals[i];
            const u64 key = topk_key(v);
            take = all ? key != 0 : (key > T || (key == T && (u32)i <= idT));
        }
        const u32 mask = __ballot_sync(FULL, take);
        if (mask) {
            const int leader = __ffs(mask) - 1;
            u32 base = 0;
            if (lane_id() == leader) base = atomicAdd(&st->out_count, (u32)__popc(mask));
            base = __shfl_sync(FULL, base, leader);
            const u32 pos = base + __popc(mask & lanemask_lt());
            if (take && pos < cap) {
                out_nodes[pos] = i;
                out_vals[pos] = v;
            }
        }
    }
}

// single-block bitonic sort of `cnt` (value desc, id asc) pairs padded to a power of two `p2`
__global__ void __launch_bounds__(1024) topk_sort_kernel(int32_t* nodes, double* vals, u32 cnt, u32 p2) {
    for (u32 i = cnt + threadIdx.x; i < p2; i += blockDim.x) { nodes[i] = 0x7fffffff; vals[i] = -1.0; }
    __syncthreads();
    for (u32 size = 2; size <= p2; size <<= 1) {
        for (u32 stride = size >> 1; stride > 0; stride >>= 1) {
            for (u32 t = threadIdx.x; t < (p2 >> 1); t += blockDim.x) {
                const u32 lo = 2 * t - (t & (stride - 1));
                const u32 hi = lo + stride;
                const bool desc_block = (lo & size) == 0; // this block sorts "first" order
                const double a = vals[lo], b = vals[hi];
                const int32_t ia = nodes[lo], ib = nodes[hi];
                const bool a_first = a > b || (a == b && ia < ib); // desired order: a before b
                if (a_first != desc_block) {
                    vals[lo] = b; vals[hi] = a;
                    nodes[lo] = ib; nodes[hi] = ia;
                }
            }
            __syncthreads();
        }
    }
}

// Select + sort on the device, results copied to host arrays of length k.
static inline cudaError_t topk_device(cudaStream_t stream, int num_sms, const double* d_vals, int32_t n, u32 k,
                                      int32_t* h_nodes, double* h_values, u64* launches, double* kth_value = nullptr,
                                      int32_t** d_nodes_out = nullptr, u32* count_out = nullptr, bool sort_on_device = false) {
    static thread_local SelectState* d_st = nullptr;
    static thread_local int32_t* d_nodes = nullptr;
    static thread_local double* d_out = nullptr;
    static thread_local u32 d_cap = 0;
    cudaError_t e;
    if (!d_st && (e = cudaMalloc((void**)&d_st, sizeof(SelectState))) != cudaSuccess) return e;
    u32 p2 = 1;
    while (p2 < k) p2 <<= 1;
    if (p2 > d_cap) {
        cudaFree(d_nodes);
        cudaFree(d_out);
        d_nodes = nullptr; d_out = nullptr; d_cap = 0;
        if ((e = cudaMalloc((void**)&d_nodes, sizeof(int32_t) * p2)) != cudaSuccess) return e;
        if ((e = cudaMalloc((void**)&d_out, sizeof(double) * p2)) != cudaSuccess) return e;
        d_cap = p2;
    }
    if ((e = cudaMemsetAsync(d_st, 0, sizeof(SelectState), stream)) != cudaSuccess) return e;
    const int gx = std::max(1, std::min(num_sms * 8, (n + 255) / 256));
    for (int shift = 56; shift >= 0; shift -= 8) {
        topk_hist_kernel<<<gx, 256, 0, stream>>>(d_vals, n, d_st, shift);
        topk_pick_kernel<<<1, 32, 0, stream>>>(d_st, shift, (u64)k);
        *launches += 2;
    }
    for (int shift = 24; shift >= 0; shift -= 8) {
        topk_idhist_kernel<<<gx, 256, 0, stream>>>(d_vals, n, d_st, shift);
        topk_idpick_kernel<<<1, 32, 0, stream>>>(d_st, shift);
        *launches += 2;
    }
    topk_collect_kernel<<<gx, 256, 0, stream>>>(d_vals, n, d_st, d_nodes, d_out, k);
    *launches += 1;
    SelectState hs;
    if ((e = cudaMemcpyAsync(&hs, d_st, sizeof(SelectState), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    const u32 cnt = std::min<u32>(hs.out_count, k);
    if (d_nodes_out) *d_nodes_out = d_nodes;
    if (count_out) *count_out = cnt;
    if (sort_on_device && cnt > 0 && !(h_nodes && h_values)) {
        u32 c2 = 1;
        while (c2 < cnt) c2 <<= 1;
        topk_sort_kernel<<<1, 1024, 0, stream>>>(d_nodes, d_out, cnt, c2);
        *launches += 1;
    }
    if (kth_value) {
        double t = 0.0;
        if (!hs.all_positive) memcpy(&t, &hs.key_T, sizeof t);
        *kth_value = t;
    }
    if (h_nodes && h_values) {
        if (cnt > 0) {
            u32 c2 = 1;
            while (c2 < cnt) c2 <<= 1;
            topk_sort_kernel<<<1, 1024, 0, stream>>>(d_nodes, d_out, cnt, c2);
            *launches += 1;
            if ((e = cudaMemcpyAsync(h_nodes, d_nodes, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
            if ((e = cudaMemcpyAsync(h_values, d_out, sizeof(double) * cnt, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
            if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
        }
        for (u32 i = cnt; i < k; ++i) { h_nodes[i] = 0; h_values[i] = 0.0; }
    }
    return cudaGetLastError();
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
__global__ void __launch_bounds__(256) ppr_bounds_kernel(int32_t n, double rsum, double pfail, double total_rw_num,
                                                          const double* __restrict__ ppr, const double* __restrict__ reserve,
                                                          double* __restrict__ upper, double* __restrict__ lower) {
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

// if_stop part 2 (algo.h:1124-1163) given the k nodes with the largest lower bounds:
//   fail[0] |= any of them has upper/lower > 1+eps;   in_topk marks them
__global__ void stop_mark_kernel(const int32_t* __restrict__ nodes, u32 k, const double* __restrict__ upper,
                                 const double* __restrict__ lower, double eps, unsigned char* __restrict__ in_topk, u32* fail) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
        const int32_t v = nodes[i];
        in_topk[v] = 1;
        if (upper[v] / lower[v] > 1.0 + eps) atomicOr(fail, 1u);
    }
}
//   fail[1] |= a node outside the top-k with ppr > 0 whose upper bound exceeds low_bound_k*(1+eps) without being
//   separated by (1+eps)/(1-eps)
__global__ void __launch_bounds__(256) stop_tail_kernel(int32_t n, const double* __restrict__ ppr, const double* __restrict__ upper,
                                                         const double* __restrict__ lower, const unsigned char* __restrict__ in_topk,
                                                         double low_bound_k, double eps, u32* fail) {
    bool bad = false;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
        if (in_topk[v] || ppr[v] <= 0) continue;
        const double u = upper[v], l = lower[v];
        if (u > low_bound_k * (1.0 + eps) && !(u > (1 + eps) / (1 - eps) * l)) bad = true;
    }
    if (__any_sync(FULL, bad) && lane_id() == 0) atomicOr(fail + 1, 1u);
}
__global__ void stop_unmark_kernel(const int32_t* __restrict__ nodes, u32 k, unsigned char* __restrict__ in_topk) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) in_topk[nodes[i]] = 0;
}

} // namespace fora
