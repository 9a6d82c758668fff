// fora_b200/csrc/common.cuh -- shared declarations for the sm_100a FORA engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "fora_b200.h"

namespace fora {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int WARP = 32;
constexpr u32 FULL = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// Graph resident in HBM.  Row offsets are kept twice: int64 for the ABI (bit-exact download) and,
// when the edge count fits, uint32 for the kernels (halves the bytes of the dependent load in the
// walk loop).  OffT selects the kernel instantiation.
// ---------------------------------------------------------------------------------------------
template <typename OffT>
struct CsrView {
    const OffT* __restrict__ ptr;    // [n+1]
    const int32_t* __restrict__ col; // [edges]
};

struct DeviceGraph {
    int32_t n = 0;
    int64_t m_decl = 0, n_edges = 0;
    bool off32 = false, has_in = false;
    int64_t* out_ptr64 = nullptr;
    u32* out_ptr32 = nullptr;
    int32_t* out_col = nullptr;
    int32_t* deg = nullptr; // out-degree
    // push kernel only: out_col with the target's out-degree packed into the id's spare high bits (engine.cu pack_columns)
    int32_t* out_colx = nullptr;
    u32 deg_shift = 0;
    int64_t* in_ptr64 = nullptr;
    u32* in_ptr32 = nullptr;
    int32_t* in_col = nullptr;
    // internal relabelling by descending in-degree (engine.cu): kernels see new ids, the ABI speaks original ids
    bool relabeled = false;
    int32_t* old2new = nullptr;
    int32_t* new2old = nullptr;
};

// ---------------------------------------------------------------------------------------------
// Counter-based RNG: Philox4x32-10 (Salmon et al., SC'11), restated from the paper.
// One call yields four 32-bit words = two walk steps (stop test + neighbour pick each).
// ---------------------------------------------------------------------------------------------
struct Philox4 {
    u32 x, y, z, w;
};
__device__ __forceinline__ Philox4 philox4x32_10(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1) {
    constexpr u32 M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const u32 hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const u32 hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const u32 n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// ---------------------------------------------------------------------------------------------
// warp helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 lanemask_lt() {
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ u32 warp_incl_scan(u32 v) {
    const int l = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(FULL, v, o);
        if (l >= o) v += t;
    }
    return v;
}
__device__ __forceinline__ u64 warp_incl_scan64(u64 v) {
    const int l = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u64 t = __shfl_up_sync(FULL, v, o);
        if (l >= o) v += t;
    }
    return v;
}

// Append `item` to list[*counter] for every lane with pred set; one atomic per warp.
// Must be called by all 32 lanes.
template <typename T>
__device__ __forceinline__ void warp_append(bool pred, T item, T* __restrict__ list, u32* counter) {
    const u32 mask = __ballot_sync(FULL, pred);
    if (mask == 0) return;
    const int leader = __ffs(mask) - 1;
    u32 base = 0;
    if (lane_id() == leader) base = atomicAdd(counter, (u32)__popc(mask));
    base = __shfl_sync(FULL, base, leader);
    if (pred) list[base + __popc(mask & lanemask_lt())] = item;
}

} // namespace fora
