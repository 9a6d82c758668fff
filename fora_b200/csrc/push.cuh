// fora_b200/csrc/push.cuh -- frontier-synchronous forward push on sm_100a.
//
// Replaces forward_local_update_linear (/root/reference/algo.h:954-1018) and its resumable
// variant forward_local_update_linear_topk (algo.h:1020-1093).  Same per-vertex rule
// (reserve += alpha*r; every out-neighbour += ((1-alpha)*r)/d_out; dangling mass -> source;
// push while residue/d_out >= rmax) on a level-synchronous schedule:
//
//   level k, phase A  every frontier vertex reads and zeroes its residue, credits its reserve
//            phase B  all scatters land as fp64 atomics; a vertex joins level k+1 exactly when one
//                     atomic moves its residue across rmax*d_out (detected from the value the
//                     atomic returns, so no per-vertex flag array and no per-level scan)
//
// One persistent cooperative kernel runs every level of every query slot of a batch; the
// frontier test and termination live on the device.  Work is edge-balanced inside a warp
// (32 frontier entries -> their concatenated edge ranges are walked 32 edges at a time, so
// column reads are coalesced and a degree-1000 vertex costs its warp 32 iterations, not 1000);
// vertices above HUB_DEG are split over the whole grid.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace fora {
namespace cg = cooperative_groups;

constexpr int PUSH_THREADS = 512;
constexpr int PUSH_WARPS = PUSH_THREADS / WARP;
constexpr int HUB_DEG = 8192;
constexpr int MAX_SLOTS = 64;

struct PushCtl {
    u32 fcount[3];   // frontier sizes, rotated by level % 3
    u32 tile_ctr[3]; // dynamic tile hand-out, rotated likewise
    u32 hub_count[2];
    u32 levels_run;
    u32 overflow;
};

struct PushArgs {
    int32_t n;
    int32_t slots;
    double alpha;
    double* reserve;  // [slots*n]   (mutated across levels: no __restrict__, read with __ldcg)
    double* residue;  // [slots*n]
    const int32_t* __restrict__ deg;
    u64* front0;                   // (slot<<32 | v)
    u64* front1;
    double* inc;      // per frontier entry: ((1-alpha)*r)/d, or (1-alpha)*r when dangling
    u32* hub;         // frontier indices of hubs of this level
    PushCtl* ctl;
    const double* __restrict__ rmax;     // [slots]
    const int32_t* __restrict__ source;  // [slots]
    u64* __restrict__ edges;             // [slots] counters
    u64* __restrict__ vertices;          // [slots]
    u64* __restrict__ levels;            // [slots]
    int32_t* __restrict__ lastlvl;       // [slots]
    u32 front_cap;
    u32 max_levels;
    u32 level_base;                      // distinguishes levels of successive launches in lastlvl
};

// One scatter: residue[slot*n+u] += inc, and detect the threshold crossing (see header).
// Called by all 32 lanes (ok = lane has an edge).
__device__ __forceinline__ void push_scatter(const PushArgs& a, bool ok, int slot, int32_t u, double inc, double rmax,
                                             u64* nxt, u32* nxt_count) {
    bool cross = false;
    if (ok) {
        const size_t g = (size_t)slot * a.n + u;
        const double old = atomicAdd(&a.residue[g], inc);
        const double nw = old + inc;
        const int32_t du = __ldg(&a.deg[u]);
        const double thr = rmax * (double)du;
        cross = du ? (old < thr && nw >= thr) : (old == 0.0);
    }
    warp_append<u64>(cross, ((u64)slot << 32) | (u32)u, nxt, nxt_count);
}

template <typename OffT>
__global__ void __launch_bounds__(PUSH_THREADS, 2) push_kernel(PushArgs a, CsrView<OffT> g) {
    cg::grid_group grid = cg::this_grid();
    __shared__ u32 s_excl[PUSH_WARPS][WARP];
    __shared__ OffT s_beg[PUSH_WARPS][WARP];
    __shared__ double s_inc[PUSH_WARPS][WARP];
    __shared__ int s_slot[PUSH_WARPS][WARP]; // slot, or ~slot when the entry is dangling
    __shared__ double s_rmax[MAX_SLOTS];
    __shared__ int32_t s_source[MAX_SLOTS];

    const int lane = lane_id();
    const int wib = threadIdx.x >> 5;
    const u32 gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 gsize = gridDim.x * blockDim.x;
    const u32 gwarp = gtid >> 5;
    PushCtl* ctl = a.ctl;

    for (int i = threadIdx.x; i < a.slots; i += blockDim.x) {
        s_rmax[i] = a.rmax[i];
        s_source[i] = a.source[i];
    }
    __syncthreads();

    u32 level = 0;
    for (;; ++level) {
        const u32 nf = *((volatile u32*)&ctl->fcount[level % 3]);
        if (nf == 0 || level >= a.max_levels) break;
        const u64* cur = (level & 1) ? a.front1 : a.front0; // written by the previous level: L2 reads only
        u64* nxt = (level & 1) ? a.front0 : a.front1;
        u32* nxt_count = &ctl->fcount[(level + 1) % 3];

        // ---------------- phase A: take the residue snapshot ----------------
        if (gtid == 0) {
            ctl->fcount[(level + 2) % 3] = 0;
            ctl->tile_ctr[(level + 1) % 3] = 0;
            ctl->hub_count[(level + 1) & 1] = 0;
            ctl->levels_run = level + 1;
        }
        for (u32 base = gwarp * WARP; base < nf; base += (gsize >> 5) * WARP) {
            const u32 i = base + lane;
            int slot = -1;
            u32 d = 0;
            if (i < nf) {
                const u64 e = __ldcg(&cur[i]);
                slot = (int)(e >> 32);
                const int32_t v = (int32_t)(u32)e;
                const size_t gi = (size_t)slot * a.n + v;
                const double r = __ldcg(&a.residue[gi]);
                a.residue[gi] = 0.0;
                a.reserve[gi] = __ldcg(&a.reserve[gi]) + r * a.alpha;
                d = (u32)__ldg(&a.deg[v]);
                a.inc[i] = d ? ((1.0 - a.alpha) * r) / (double)d : r * (1.0 - a.alpha);
            }
            // per-slot work counters (cost model of --balanced, roofline accounting)
            const int slot0 = __shfl_sync(FULL, slot, 0);
            if (__all_sync(FULL, slot == slot0 || slot < 0)) {
                const u32 dsum = warp_sum(d);
                const u32 cnt = __popc(__ballot_sync(FULL, slot >= 0));
                if (lane == 0) {
                    atomicAdd(&a.edges[slot0], (u64)dsum);
                    atomicAdd(&a.vertices[slot0], (u64)cnt);
                    if (atomicMax(&a.lastlvl[slot0], (int)(a.level_base + level + 1)) < (int)(a.level_base + level + 1))
                        atomicAdd(&a.levels[slot0], 1ull);
                }
            } else if (slot >= 0) {
                atomicAdd(&a.edges[slot], (u64)d);
                atomicAdd(&a.vertices[slot], 1ull);
                if (atomicMax(&a.lastlvl[slot], (int)(a.level_base + level + 1)) < (int)(a.level_base + level + 1))
                    atomicAdd(&a.levels[slot], 1ull);
            }
        }
        grid.sync();

        // ---------------- phase B: scatter, 32 frontier entries per warp tile ----------------
        const u32 ntiles = (nf + WARP - 1) / WARP;
        u32* tile_ctr = &ctl->tile_ctr[level % 3];
        u32* hub_count = &ctl->hub_count[level & 1];
        // first tile of every warp is static (no atomic at all while the frontier is smaller than the
        // grid: a same-address atomic per warp per level costs more than the level itself); further
        // tiles are handed out dynamically
        const u32 nwarps = gsize >> 5;
        for (u32 round = 0;; ++round) {
            u32 t = gwarp;
            if (round > 0) {
                if (ntiles <= nwarps) break;
                if (lane == 0) t = nwarps + atomicAdd(tile_ctr, 1u);
                t = __shfl_sync(FULL, t, 0);
            }
            if (t >= ntiles) break;
            const u32 i = t * WARP + lane;
            u32 d = 0;
            if (i < nf) {
                const u64 e = __ldcg(&cur[i]);
                const int slot = (int)(e >> 32);
                const int32_t v = (int32_t)(u32)e;
                const OffT beg = g.ptr[v];
                const u32 dreal = (u32)(g.ptr[v + 1] - beg);
                s_beg[wib][lane] = beg;
                s_inc[wib][lane] = __ldcg(&a.inc[i]);
                s_slot[wib][lane] = dreal ? slot : ~slot;
                d = dreal ? dreal : 1u;
                if (dreal > (u32)HUB_DEG) {
                    a.hub[atomicAdd(hub_count, 1u)] = i;
                    d = 0;
                }
            }
            const u32 incl = warp_incl_scan(d);
            s_excl[wib][lane] = incl - d;
            const u32 total = __shfl_sync(FULL, incl, 31);
            __syncwarp();
            for (u32 e0 = 0; e0 < total; e0 += WARP) {
                const u32 ee = e0 + lane;
                const bool ok = ee < total;
                int slot = 0;
                int32_t u = 0;
                double inc = 0.0, rmax = 0.0;
                if (ok) {
                    // last j with excl[j] <= ee (entries of width 0 share their successor's offset)
                    int lo = 0;
#pragma unroll
                    for (int step = 16; step > 0; step >>= 1)
                        if (lo + step < WARP && s_excl[wib][lo + step] <= ee) lo += step;
                    const int sj = s_slot[wib][lo];
                    slot = sj < 0 ? ~sj : sj;
                    inc = s_inc[wib][lo];
                    rmax = s_rmax[slot];
                    u = sj < 0 ? s_source[slot] : __ldg(&g.col[s_beg[wib][lo] + (OffT)(ee - s_excl[wib][lo])]);
                }
                push_scatter(a, ok, slot, u, inc, rmax, nxt, nxt_count);
            }
            __syncwarp();
        }
        grid.sync();

        // ---------------- phase B2: hubs, edges split over the whole grid ----------------
        const u32 nh = *((volatile u32*)hub_count);
        if (nh > 0) {
            for (u32 h = 0; h < nh; ++h) {
                const u32 i = __ldcg(&a.hub[h]);
                const u64 e = __ldcg(&cur[i]);
                const int slot = (int)(e >> 32);
                const int32_t v = (int32_t)(u32)e;
                const OffT beg = g.ptr[v];
                const u32 d = (u32)(g.ptr[v + 1] - beg);
                const double inc = __ldcg(&a.inc[i]);
                const double rmax = s_rmax[slot];
                for (u32 e0 = gwarp * WARP; e0 < d; e0 += (gsize >> 5) * WARP) {
                    const u32 ee = e0 + lane;
                    const bool ok = ee < d;
                    const int32_t u = ok ? __ldg(&g.col[beg + (OffT)ee]) : 0;
                    push_scatter(a, ok, slot, u, inc, rmax, nxt, nxt_count);
                }
            }
            grid.sync();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Slot initialisation: residue[s] = 1 (algo.h:976 / query.h:856); a source with no out-edges
// keeps everything as reserve (algo.h:961-965).  seed_source=1 also makes {s} the level-0
// frontier unconditionally (algo.h:973).
// ---------------------------------------------------------------------------------------------
__global__ void push_init_kernel(int32_t n, int32_t slots, const int32_t* __restrict__ source,
                                 const int32_t* __restrict__ deg, double* __restrict__ reserve,
                                 double* __restrict__ residue, u64* __restrict__ front0, PushCtl* ctl, int seed_source,
                                 int32_t* __restrict__ slot_state) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= slots) return;
    const int32_t s = source[slot];
    if (s < 0) { // unused slot of a partial wave
        slot_state[slot] = 0;
        return;
    }
    const size_t gi = (size_t)slot * n + s;
    if (deg[s] == 0) {
        reserve[gi] = 1.0;
        slot_state[slot] = 2; // dangling source: nothing to push, rsum = 0
    } else {
        residue[gi] = 1.0;
        slot_state[slot] = 1;
        if (seed_source) front0[atomicAdd(&ctl->fcount[0], 1u)] = ((u64)slot << 32) | (u32)s;
    }
}

// Level-0 frontier of a resumable round: every vertex of an active slot with
// residue/d_out >= rmax (dangling vertices: any positive residue, x/0 = +inf in algo.h:1039).
__global__ void __launch_bounds__(256) push_seed_kernel(int32_t n, const int32_t* __restrict__ deg,
                                                         const double* __restrict__ residue,
                                                         const double* __restrict__ rmax,
                                                         const int32_t* __restrict__ slot_active,
                                                         u64* __restrict__ front0, PushCtl* ctl) {
    const int slot = blockIdx.y;
    if (!slot_active[slot]) return;
    const double rm = rmax[slot];
    const double* __restrict__ res = residue + (size_t)slot * n;
    const int nwarp_iters = (n + WARP - 1) / WARP;
    for (int wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < nwarp_iters; wi += (gridDim.x * blockDim.x) >> 5) {
        const int v = wi * WARP + lane_id();
        bool pred = false;
        if (v < n) {
            const double r = res[v];
            const int32_t d = deg[v];
            pred = d ? (r >= rm * (double)d) : (r > 0.0);
        }
        warp_append<u64>(pred, ((u64)slot << 32) | (u32)v, front0, &ctl->fcount[0]);
    }
}

// ---------------------------------------------------------------------------------------------
// Deterministic per-slot reduction of the residue vector: rsum (fixed summation tree) and the
// number of non-zero entries.  grid = (blocks, slots).
// ---------------------------------------------------------------------------------------------
constexpr int RED_THREADS = 256;
__global__ void __launch_bounds__(RED_THREADS) residue_partial_kernel(int32_t n, const double* __restrict__ residue,
                                                                      double* __restrict__ part_sum,
                                                                      u32* __restrict__ part_nnz) {
    __shared__ double s_sum[RED_THREADS];
    __shared__ u32 s_nnz[RED_THREADS];
    const int slot = blockIdx.y;
    const double* __restrict__ res = residue + (size_t)slot * n;
    const int per_block = (n + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per_block, hi = min(n, lo + per_block);
    double s = 0.0;
    u32 c = 0;
    for (int v = lo + threadIdx.x; v < hi; v += RED_THREADS) {
        const double r = res[v];
        s += r;
        c += r > 0.0;
    }
    s_sum[threadIdx.x] = s;
    s_nnz[threadIdx.x] = c;
    __syncthreads();
    for (int o = RED_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
            s_nnz[threadIdx.x] += s_nnz[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        part_sum[(size_t)slot * gridDim.x + blockIdx.x] = s_sum[0];
        part_nnz[(size_t)slot * gridDim.x + blockIdx.x] = s_nnz[0];
    }
}
__global__ void residue_final_kernel(int nblocks, const double* __restrict__ part_sum, const u32* __restrict__ part_nnz,
                                     double* __restrict__ rsum, u64* __restrict__ nnz) {
    const int slot = blockIdx.x;
    if (threadIdx.x != 0) return;
    double s = 0.0;
    u64 c = 0;
    for (int b = 0; b < nblocks; ++b) {
        s += part_sum[(size_t)slot * nblocks + b];
        c += part_nnz[(size_t)slot * nblocks + b];
    }
    rsum[slot] = s;
    nnz[slot] = c;
}

} // namespace fora
