// fora_b200/csrc/push.cuh -- frontier-synchronous forward push on sm_100a.
//
// Replaces forward_local_update_linear (/root/reference/algo.h:954-1018) and its resumable
// variant forward_local_update_linear_topk (algo.h:1020-1093).  Same per-vertex rule
// (reserve += alpha*r; every out-neighbour += ((1-alpha)*r)/d_out; dangling mass -> source;
// push while residue/d_out >= rmax) on a level-synchronous schedule:
//
//   level k, phase A  every frontier vertex reads and zeroes its residue (one exchange) and logs its reserve credit
//                     (vertex, r); apply_log_kernel adds alpha*r to the reserve vectors once per wave
//            phase B  all scatters land as fp64 atomics; a vertex joins level k+1 exactly when one
//                     atomic moves its residue across rmax*d_out (detected from the value the
//                     atomic returns, so no per-vertex flag array and no per-level scan)
//
// One persistent cooperative kernel runs every level of every query slot of a batch; the
// frontier test and termination live on the device.  Work is edge-balanced across the grid: the
// edges of a level are laid on one slot-major line (exclusive scan of the frontier's out-degrees) and
// cut into tiles that the CTAs take round-robin, so a power-law hub is simply split across CTAs, no
// warp waits for a straggler, and the grid sweeps the slots front to back (all slots share each level's two
// barriers; with tens of slots their vectors do not stay L2-resident between levels -- the kernel runs at the
// memory system's random read-modify-write rate).
#pragma once
#include <cooperative_groups.h>
#include <cstddef>

#include "common.cuh"

namespace fora {
namespace cg = cooperative_groups;

constexpr int PUSH_THREADS = 512;
constexpr int PUSH_WARPS = PUSH_THREADS / WARP;
constexpr int MAX_SLOTS = 64; // also bounded by the 8 slot bits of a frontier entry
constexpr int MAX_PUSH_CTAS = 512;
constexpr u32 TILE_MIN = 2048;   // smallest edge range worth giving to a CTA
constexpr u32 TILE_MAX = 32768;  // tiles of a large level: the grid sweeps the edge line 296*32K edges at a time (16K: -2.8 % edges/s, 8K: -7 %)
#ifndef CFG_PUSH_UA
#define CFG_PUSH_UA 1
#endif
constexpr int PUSH_UA = CFG_PUSH_UA;       // frontier entries per thread per phase-A batch (measured 1: 19.2, 2: 18.9, 4: 18.5, 8: 16.6 G edges/s)
#ifndef CFG_PUSH_BATCH
#define CFG_PUSH_BATCH 1024
#endif
constexpr int PUSH_BATCH = CFG_PUSH_BATCH; // frontier entries staged in shared memory per phase-B batch
#ifndef CFG_PUSH_UB
#define CFG_PUSH_UB 2
#endif
constexpr int PUSH_UB = CFG_PUSH_UB;       // edges in flight per lane in phase B
#ifndef CFG_PUSH_WQ
#define CFG_PUSH_WQ 256
#endif
constexpr int PUSH_WQ = CFG_PUSH_WQ;     // per-warp queue of crossing vertices
constexpr int SCAN_K = 32;               // dense scan: vertices per thread and tile (tile = 32 * 512 vertices)
constexpr u32 SCAN_HCAP = 4096;          // hits of a tile handled per batch (their lists alias the idle phase-B staging arrays)

// frontier entry: [slot:8][min(out-degree, 2^24-1):24][vertex:32].  Whoever appends a vertex has just loaded its
// out-degree for the threshold test, so carrying it saves phase A one random access per vertex.
constexpr u32 DEG_SAT = 0xffffffu;
__host__ __device__ __forceinline__ u64 make_entry(int slot, u32 deg, int32_t v) {
    return ((u64)(u32)slot << 56) | ((u64)(deg < DEG_SAT ? deg : DEG_SAT) << 32) | (u32)v;
}
__host__ __device__ __forceinline__ int entry_slot(u64 e) { return (int)(e >> 56); }
__host__ __device__ __forceinline__ u32 entry_deg24(u64 e) { return (u32)(e >> 32) & DEG_SAT; }
__host__ __device__ __forceinline__ int32_t entry_vertex(u64 e) { return (int32_t)(u32)e; }

struct PushCtl {
    u32 fcount[3][MAX_SLOTS]; // per-slot frontier sizes: push_kernel rotates by level % 3; push2 / tail kernels index [buffer][slot]
    u32 levels_run;
    u32 pad[3];
    u32 par[MAX_SLOTS];       // push2 / tail kernels: which buffer (front0 / front1) holds the slot's current frontier
};

struct PushArgs {
    int32_t n;
    int32_t slots;
    double alpha;
    double* reserve;  // [slots*n]   (mutated across levels: no __restrict__, read with __ldcg)
    double* residue;  // [slots*n]
    const int32_t* __restrict__ deg;
    u64* front0;      // [slots*n]: slot s owns [s*n, (s+1)*n); entries make_entry(slot, deg, v)
    u64* front1;
    double* inc;      // per frontier entry (global index): ((1-alpha)*r)/d, or (1-alpha)*r when dangling
    u32* eoff;        // per frontier entry: edge offset inside its CTA chunk
    u64* block_sum;   // [gridDim.x] edges per CTA chunk of the current level
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
    u64* trace;                          // optional [4*trace_cap]: per level {t0 ns, nf, E, t after phase A}
    u32 trace_cap;
    u32 tile_max;                        // edges per tile of a large level
    u32 l2_hints;                        // 1: evict_last on residue atomics / degree loads, evict_first on streams
    const int32_t* __restrict__ colx;    // optional packed columns: id | min(d_out(id), dmax) << deg_shift (null: plain g.col)
    u32 deg_shift;                       // id bits of a packed column entry; the remaining high bits hold the out-degree code
    int32_t* err;                        // [1] set when the level cap stops a launch with a non-empty frontier
    // reserve credit log: phase A appends (vertex, residue pushed) instead of the random read-modify-write of reserve[v];
    // apply_log_kernel adds alpha*r slot by slot later (engine.cu).  Entries that do not fit take the direct update.
    int32_t* log_v;                      // [slots*log_cap], null: no log
    double* log_r;                       // [slots*log_cap]
    u32 log_cap;
    u32* log_cur;                        // [MAX_SLOTS] entries logged so far in this wave, per slot
    // dense slot-levels: a slot whose frontier has at least dense_min entries scatters with RED (no return value) and its next
    // frontier is found by one scan of its residue vector, which also does that frontier's phase A (0xffffffff: never)
    u32 dense_min;
    // lockstep phase B: the edge line is worked through in groups of whole slots of at least this many edges, with a grid barrier
    // between groups, so that the grid never has more than one or two residue vectors live in the L2 (0: one sweep, no barriers)
    u64 lockstep_edges;
    // edge lists of dense slot-levels: the scan that finds a slot's next frontier also gathers that frontier's column words into a
    // sequential (target, index of the pushing entry) list, and the next level's scatters of the slot stream that list (push_el_adds)
    // instead of going through phase B's staging / owner search.  el = null: off.
    uint4* el;          // [2][slots][el_cap] by level parity: (target, 0, increment as two words) -- the add pass needs ONE load per edge
    u32 el_cap;         // edges per slot and level; a level that does not fit falls back to the tiles
    u32* el_count;      // [2][MAX_SLOTS] edges listed for the slot's frontier of a level of that parity
    u32* el_bad;        // [2][MAX_SLOTS] that list overflowed
    u32 el_prefetch;    // vertices at the front of the slot's residue vector brought into the L2 before its adds start (hot after the relabelling)
    u32 debug_el_skip;  // ablation: adds to vertices below this id are dropped
    u32 debug_el;       // ablation (wrong answers, timing only): 1 adds go to hashed uniform targets, 2 no adds at all, 4 no hot accumulators
};

// dynamic shared memory of the push kernel (~70 KB, two CTAs per SM)
template <typename OffT>
struct PushSmem {
    u64 base[MAX_PUSH_CTAS + 1]; // exclusive prefix of block_sum
    u64 G[PUSH_BATCH + 1];       // global edge offset of each entry of the current batch
    double inc[PUSH_BATCH];
    OffT beg[PUSH_BATCH];
    int slot[PUSH_BATCH];        // slot, or ~slot when the entry is dangling
    u64 wqueue[PUSH_WARPS][PUSH_WQ]; // crossing vertices, one private queue per warp (no shared atomics)
    u64 warp_tot[2][PUSH_WARPS]; // double-buffered by batch parity: one CTA barrier per scan
    double rmax[MAX_SLOTS];
    int32_t source[MAX_SLOTS];
    u32 cnt_edges[MAX_SLOTS], cnt_verts[MAX_SLOTS];
    u32 fbase[MAX_SLOTS + 1];    // the level's frontier = the slots' segments concatenated in slot order
    u32 logbase[MAX_SLOTS];      // log position of the first frontier entry of each slot at this level
    u32 prevcnt[MAX_SLOTS];      // frontier size of each slot at this level (advances logbase at the next one)
    u32 i0;
    u32 step_ctr;                // dynamic hand-out of 32*PUSH_UB-edge steps inside a batch
    // dense slot-levels (push_dense_scan)
    unsigned char dense[MAX_SLOTS];  // this level: the slot scatters with RED and is scanned afterwards
    unsigned char adone[MAX_SLOTS];  // this level: the slot's frontier was made by a scan, its phase A is done
    u32 ndense;
    u32 nf_all;                      // this level: some slot has a frontier
    unsigned char el_ok[MAX_SLOTS];  // this level: the slot's edges are in its edge list (no phase A / phase B entries for it)
    u32 sc_cnt[SCAN_K * PUSH_WARPS];
    u32 sc_wtot[PUSH_WARPS];
    u32 sc_base, sc_ebase, sc_etot;
    u64 gslot[MAX_SLOTS + 1];           // lockstep phase B: where each slot starts on the level's edge line
};


// L2 eviction-priority hints (sm_80+ createpolicy / .L2::cache_hint): the vectors every edge touches at
// random (residue, out-degree) are kept with evict_last, everything streamed once per level goes through
// with evict_first -- pinning without giving up a fixed carve-out of the L2.
__device__ __forceinline__ u64 l2_policy_evict_last() {
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ u64 l2_policy_evict_first() {
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ double atomic_add_f64_hint(double* addr, double v, u64 policy) {
    double old;
    asm volatile("atom.global.add.L2::cache_hint.f64 %0, [%1], %2, %3;" : "=d"(old) : "l"(addr), "d"(v), "l"(policy) : "memory");
    return old;
}
__device__ __forceinline__ void red_add_f64_hint(double* addr, double v, u64 policy) {
    asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(addr), "d"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ int32_t ld_s32_hint(const int32_t* addr, u64 policy) {
    int32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(addr), "l"(policy));
    return v;
}
__device__ __forceinline__ int32_t ld_col_stream(const int32_t* addr, u64 policy) {
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(addr), "l"(policy));
    return v;
}

// block-wide inclusive scan of one u64 per thread; returns the inclusive value, *total = block sum.  One CTA
// barrier per call: the warp totals go to buffer `buf` (callers alternate 0/1), every warp scans the 16 totals itself,
// and the next call writes the other buffer, so a slow reader of this one is never overtaken.
template <typename OffT>
__device__ __forceinline__ u64 block_incl_scan(PushSmem<OffT>& sm, u64 v, u64* total, int buf) {
    const int lane = lane_id(), w = threadIdx.x >> 5;
    u64 incl = warp_incl_scan64(v);
    if (lane == 31) sm.warp_tot[buf][w] = incl;
    __syncthreads();
    const u64 t = lane < PUSH_WARPS ? sm.warp_tot[buf][lane] : 0;
    const u64 ti = warp_incl_scan64(t); // inclusive over warps
    *total = __shfl_sync(FULL, ti, PUSH_WARPS - 1);
    const u64 before = __shfl_sync(FULL, ti, w > 0 ? w - 1 : 0);
    if (w > 0) incl += before;
    return incl;
}

// frontier entry i of the level (slot-major concatenation of the per-slot segments)
template <typename OffT>
__device__ __forceinline__ u64 frontier_entry(const PushArgs& a, const PushSmem<OffT>& sm, const u64* cur, u32 i) {
    int lo = 0, hi = a.slots; // last slot with fbase[slot] <= i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (sm.fbase[mid] <= i) lo = mid;
        else hi = mid;
    }
    return __ldcs(&cur[(size_t)lo * a.n + (i - sm.fbase[lo])]);
}

template <typename OffT>
__device__ __forceinline__ void push_flush_counters(const PushArgs& a, PushSmem<OffT>& sm) {
    if (threadIdx.x < (u32)a.slots) {
        const int sl = threadIdx.x;
        const u32 v = sm.cnt_verts[sl];
        if (v) {
            atomicAdd(&a.edges[sl], (u64)sm.cnt_edges[sl]);
            atomicAdd(&a.vertices[sl], (u64)v);
            sm.cnt_edges[sl] = 0;
            sm.cnt_verts[sl] = 0;
        }
    }
}

// ---- phase A: CTA `rank` of `count` takes a contiguous chunk of the frontier: snapshot + zero the
// residues, credit the reserves, and lay the chunk's edges out on a line (eoff = exclusive scan
// of the out-degrees; a dangling vertex owns one pseudo-edge back to its slot's source).
// Every thread owns PUSH_UA consecutive entries and issues all their loads before using any.
template <typename OffT, bool DN>
__device__ __forceinline__ void push_phase_a(const PushArgs& a, PushSmem<OffT>& sm, const u64* cur, u32 nf, u32 level,
                                             u32 rank, u32 count) {
    const int lane = lane_id();
    const u32 cs = (nf + count - 1) / count;
    const u32 lo_i = min(nf, rank * cs), hi_i = min(nf, lo_i + cs);
    u64 carry = 0;
    int parity = 0;
    for (u32 b0 = lo_i; b0 < hi_i; b0 += PUSH_THREADS * PUSH_UA) {
        const u32 i0 = b0 + threadIdx.x * PUSH_UA;
        u64 e[PUSH_UA];
        double r[PUSH_UA], rs[PUSH_UA];
        u32 d[PUSH_UA], lp[PUSH_UA];
#pragma unroll
        for (int k = 0; k < PUSH_UA; ++k) e[k] = (i0 + k < hi_i) ? frontier_entry(a, sm, cur, i0 + k) : ~0ull;
        bool ad[PUSH_UA]; // the entry's phase A was done by the scan that made it: only its edge offset is missing
#pragma unroll
        for (int k = 0; k < PUSH_UA; ++k) {
            r[k] = rs[k] = 0.0;
            d[k] = 0;
            ad[k] = false;
            if (e[k] != ~0ull) {
                const int slot = entry_slot(e[k]);
                const size_t gi = (size_t)slot * a.n + (u32)e[k];
                ad[k] = DN && sm.adone[slot];
                if (!ad[k]) {
                    // read + zero in ONE L2 operation.  (A load followed by a plain `= 0.0` store is a trap: when the compiler hoists
                    // the store right behind the load, it reaches the L2 while the sector's fill is still pending and takes a slow
                    // path -- phase A ran 3x slower.  Ordering the store behind the load by a data dependency fixed that; the
                    // exchange is 1.4 % faster still.)
                    r[k] = __longlong_as_double((long long)atomicExch((unsigned long long*)&a.residue[gi], 0ull));
                    lp[k] = sm.logbase[slot] + (i0 + k - sm.fbase[slot]);
                    if (!(a.log_v && lp[k] < a.log_cap)) {
                        lp[k] = 0xffffffffu; // no room (or no log): direct update of the reserve
                        rs[k] = __ldcg(&a.reserve[gi]);
                    }
                }
                d[k] = entry_deg24(e[k]);
                if (d[k] == DEG_SAT) d[k] = (u32)__ldg(&a.deg[(u32)e[k]]);
            }
        }
        u32 esum = 0, vcnt = 0, dsum = 0;
        u32 loc[PUSH_UA];
        int slot_first = -1;
        bool same = true;
#pragma unroll
        for (int k = 0; k < PUSH_UA; ++k) {
            if (e[k] == ~0ull) continue;
            const int slot = entry_slot(e[k]);
            const size_t gi = (size_t)slot * a.n + (u32)e[k];
            if (!ad[k]) {
                if (lp[k] != 0xffffffffu) {
                    const size_t li = (size_t)slot * a.log_cap + lp[k];
                    a.log_v[li] = (int32_t)(u32)e[k];
                    a.log_r[li] = r[k];
                } else {
                    a.reserve[gi] = rs[k] + r[k] * a.alpha;
                }
                // increments are kept per slot segment (index inside the slot's frontier), so that a scan can write them for a
                // frontier whose place in the next level's slot-major order is not known yet
                a.inc[(size_t)slot * a.n + (i0 + k - sm.fbase[slot])] = d[k] ? ((1.0 - a.alpha) * r[k]) / (double)d[k] : r[k] * (1.0 - a.alpha);
                dsum += d[k];
                ++vcnt;
            }
            loc[k] = esum; // thread-local exclusive offset, completed below
            esum += d[k] ? d[k] : 1u;
            if (slot_first < 0) slot_first = slot;
            else same = same && slot == slot_first;
        }
        // per-slot work counters (cost model of --balanced, roofline accounting): shared memory first
        const int slot0 = __shfl_sync(FULL, slot_first, 0);
        if (__all_sync(FULL, same && (slot_first == slot0 || slot_first < 0))) {
            const u32 ds = warp_sum(dsum), vc = warp_sum(vcnt);
            if (lane == 0 && slot0 >= 0 && vc) {
                atomicAdd(&sm.cnt_edges[slot0], ds);
                atomicAdd(&sm.cnt_verts[slot0], vc);
            }
        } else {
#pragma unroll
            for (int k = 0; k < PUSH_UA; ++k)
                if (e[k] != ~0ull && !ad[k]) {
                    atomicAdd(&sm.cnt_edges[entry_slot(e[k])], d[k]);
                    atomicAdd(&sm.cnt_verts[entry_slot(e[k])], 1u);
                }
        }
        u64 total;
        const u64 incl = block_incl_scan(sm, (u64)esum, &total, parity);
        parity ^= 1;
        const u32 tbase = (u32)(carry + incl - esum);
#pragma unroll
        for (int k = 0; k < PUSH_UA; ++k)
            if (e[k] != ~0ull) a.eoff[i0 + k] = loc[k] + tbase;
        carry += total;
    }
    if (threadIdx.x == 0) a.block_sum[rank] = carry;
    // flush the per-slot counters of this CTA once all its warps are past their last shared atomics (a dense scan of the previous
    // level may have left some too)
    __syncthreads();
    push_flush_counters(a, sm);
}

// append this warp's queue (entries of ONE slot) to that slot's segment of the next frontier
__device__ __forceinline__ void push_flush_warp(const PushArgs& a, const u64* myq, u32 wq, int wq_slot, u64* nxt, u32* nxt_count) {
    const int lane = lane_id();
    u32 base = 0;
    __syncwarp();
    if (lane == 0) base = atomicAdd(&nxt_count[wq_slot], wq);
    base = __shfl_sync(FULL, base, 0);
    u64* dst = nxt + (size_t)wq_slot * a.n + base;
    for (u32 t = lane; t < wq; t += WARP) dst[t] = myq[t];
    __syncwarp();
}

// ---- phase B: the level's edges form one line of length E (chunk bases + eoff), slot-major because
// the frontier is.  The line is cut into tiles; CTA `rank` takes tiles rank, rank+count, ...: work is
// perfectly edge-balanced whatever the degree distribution (a hub is simply cut across tiles) and the
// whole grid sweeps the line front to back together, so at any moment it touches the residue vectors
// of one or two slots only -- they stay L2-resident although a wave holds many slots.  Entries are staged
// PUSH_BATCH at a time in shared memory; inside a batch the warps run independently (no CTA barrier in
// the edge loop): a warp step covers 32*PUSH_UB consecutive edges, every lane keeps PUSH_UB scatters in
// flight, owners are found by binary search in shared memory, column reads are coalesced and streamed,
// and crossing vertices collect in a private per-warp queue that is appended to the slot's next
// frontier with one global atomic per flush.
template <typename OffT, bool DN>
__device__ __forceinline__ void push_phase_b(const PushArgs& a, const CsrView<OffT>& g, PushSmem<OffT>& sm, const u64* cur,
                                             u32 nf, u32 level, u32 rank, u32 count, u64* nxt, u32* nxt_count) {
    const int lane = lane_id(), w = threadIdx.x >> 5;
    const u32 cs = (nf + count - 1) / count;
    for (u32 b = threadIdx.x; b < count; b += PUSH_THREADS) sm.base[b] = __ldcg(&a.block_sum[b]); // one parallel round
    __syncthreads();
    if (w == 0) {
        u64 carry = 0;
        for (u32 b0 = 0; b0 < count; b0 += WARP) {
            const u32 b = b0 + lane;
            const u64 v = b < count ? sm.base[b] : 0;
            const u64 incl = warp_incl_scan64(v);
            if (b < count) sm.base[b] = carry + incl - v;
            carry += __shfl_sync(FULL, incl, 31);
        }
        if (lane == 0) sm.base[count] = carry;
    }
    __syncthreads();
    const u64 E = sm.base[count];
    if (a.trace && rank == 0 && threadIdx.x == 0 && level < a.trace_cap) a.trace[4 * level + 2] = E;
    if (E == 0) return;
    if (a.lockstep_edges) {
        for (u32 t = threadIdx.x; t <= (u32)a.slots; t += PUSH_THREADS) {
            const u32 i = sm.fbase[t];
            sm.gslot[t] = i < nf ? sm.base[i / cs] + __ldcg(&a.eoff[i]) : E;
        }
        __syncthreads();
    }
    const u64 pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
    u32 wq = 0;       // entries in this warp's queue (warp-uniform register)
    int wq_slot = 0;  // the slot they belong to
    u64* myq = sm.wqueue[w];
    // (Rejected after measurement: tiles claimed dynamically from a cursor with guided, shrinking sizes -- 16.7 instead
    // of 18.4 G edges/s; a tile start costs a 32-ary search in global memory plus a full batch staging, so fewer,
    // larger static tiles win although the grid barrier then waits for the CTAs that drew one tile more.)
    int gs = 0;
    for (u64 g_lo = 0; g_lo < E;) {
    u64 g_hi = E;
    if (a.lockstep_edges) { // the group ends at the first slot boundary that leaves it at least lockstep_edges edges
        while (gs < a.slots && sm.gslot[gs] <= g_lo) ++gs; // first slot that starts behind g_lo
        while (gs < a.slots && sm.gslot[gs] - g_lo < a.lockstep_edges) ++gs;
        if (gs < a.slots) g_hi = sm.gslot[gs];
    }
    u64 T = (g_hi - g_lo + count - 1) / count;
    T = T < TILE_MIN ? TILE_MIN : (T > a.tile_max ? a.tile_max : T);
    for (u64 lo = g_lo + (u64)rank * T; lo < g_hi; lo += (u64)count * T) {
        const u64 hi = min(g_hi, lo + T);
        // first entry of the tile: largest i with G(i) <= lo, G(i) = base[i / cs] + eoff[i]
        if (w == 0) {
            u32 c = 0; // largest chunk with base[c] <= lo (only trailing chunks are empty, and they sit at E)
            {
                u32 l = 0, h = count;
                while (h - l > 1) {
                    const u32 mid = (l + h) >> 1;
                    if (sm.base[mid] <= lo) l = mid;
                    else h = mid;
                }
                c = l;
            }
            const u64 target = lo - sm.base[c];
            u32 l = min(nf, c * cs), h = min(nf, l + cs); // answer in [l, h), eoff[l] == 0 <= target
            while (h - l > 1) {                            // 32-ary search
                const u32 span = h - l;
                const u32 step = (span + WARP - 1) / WARP;
                const u32 pos = l + lane * step;
                const bool le = pos < h && (u64)__ldcg(&a.eoff[pos]) <= target;
                const u32 m = __ballot_sync(FULL, le);
                const int last = 31 - __clz(m); // lane 0 always satisfies
                const u32 nl = l + last * step;
                h = min(h, nl + step);
                l = nl;
            }
            if (lane == 0) sm.i0 = l;
        }
        __syncthreads();
        u32 i_cur = sm.i0;
        for (;;) {
            const u32 cnt = min((u32)PUSH_BATCH, nf - i_cur);
            for (u32 t = threadIdx.x; t <= cnt; t += PUSH_THREADS) {
                const u32 i = i_cur + t;
                if (t < cnt) {
                    const u64 e = frontier_entry(a, sm, cur, i);
                    const int slot = entry_slot(e);
                    const int32_t v = entry_vertex(e);
                    const OffT beg = g.ptr[v];
                    const bool dang = entry_deg24(e) == 0;
                    sm.G[t] = sm.base[i / cs] + __ldcg(&a.eoff[i]);
                    sm.beg[t] = beg;
                    sm.inc[t] = __ldcg(&a.inc[(size_t)slot * a.n + (i - sm.fbase[slot])]);
                    sm.slot[t] = dang ? ~slot : slot;
                } else {
                    sm.G[cnt] = i < nf ? sm.base[i / cs] + __ldcg(&a.eoff[i]) : E;
                    sm.step_ctr = PUSH_WARPS; // steps 0..PUSH_WARPS-1 are pre-assigned, one per warp
                }
            }
            __syncthreads();
            const u64 x_lo = max(lo, sm.G[0]), x_hi = min(hi, sm.G[cnt]);
            // warp w starts with step w, further steps are handed out dynamically (no warp waits at the batch barrier
            // for a neighbour that drew one more step)
            for (u64 step = w;; ) {
                const u64 xb = x_lo + step * (u64)(WARP * PUSH_UB);
                if (xb >= x_hi) break;
                if (wq > PUSH_WQ - WARP * PUSH_UB) { // make room: one global atomic per flush
                    push_flush_warp(a, myq, wq, wq_slot, nxt, nxt_count);
                    wq = 0;
                }
                int slot[PUSH_UB];
                int32_t u[PUSH_UB];
                u32 dcode[PUSH_UB]; // packed columns: the target's out-degree came along with its id (dmax: too large, load it)
                double inc[PUSH_UB], old[PUSH_UB];
                bool ok[PUSH_UB];
                const u32 dmax = a.colx ? (0xffffffffu >> a.deg_shift) : 0u, idmask = a.colx ? ((1u << a.deg_shift) - 1u) : 0xffffffffu;
                const int32_t* __restrict__ colp = a.colx ? a.colx : g.col;
                u32 l = 0;
#pragma unroll
                for (int k = 0; k < PUSH_UB; ++k) {
                    const u64 x = xb + (u64)k * WARP + lane;
                    ok[k] = x < x_hi;
                    slot[k] = 0; u[k] = 0; inc[k] = 0.0; dcode[k] = dmax;
                    if (ok[k]) {
                        // largest t in [l, h) with G[t] <= x; after the first edge the owner moves forward by at
                        // most 32 entries (every entry owns >= 1 edge)
                        u32 h = k == 0 ? cnt : min(cnt, l + WARP + 1);
                        while (h - l > 1) {
                            const u32 mid = (l + h) >> 1;
                            if (sm.G[mid] <= x) l = mid;
                            else h = mid;
                        }
                        const int sj = sm.slot[l];
                        slot[k] = sj < 0 ? ~sj : sj;
                        inc[k] = sm.inc[l];
                        if (sj < 0) {
                            u[k] = sm.source[slot[k]];
                        } else {
                            const int32_t* cp = &colp[sm.beg[l] + (OffT)(x - sm.G[l])];
                            const u32 raw = (u32)(a.l2_hints ? ld_col_stream(cp, pol_stream) : __ldcs(cp));
                            u[k] = (int32_t)(raw & idmask);
                            if (a.colx) dcode[k] = raw >> a.deg_shift;
                        }
                    }
                }
                // the out-degrees of the targets are fetched alongside the atomics (both depend only on the
                // column values), so a step is two dependent memory round trips instead of three; with packed
                // columns the degree is already there and the random load disappears for all but the largest hubs
                int32_t du[PUSH_UB];
                bool dn[PUSH_UB]; // dense slot-level: RED, nobody waits for the old value, the scan finds the next frontier
#pragma unroll
                for (int k = 0; k < PUSH_UB; ++k) {
                    du[k] = 0;
                    dn[k] = false;
                    if (ok[k]) {
                        double* rp = &a.residue[(size_t)slot[k] * a.n + u[k]];
                        dn[k] = DN && sm.dense[slot[k]];
                        if (dn[k]) {
                            if (a.l2_hints) red_add_f64_hint(rp, inc[k], pol_keep);
                            else atomicAdd(rp, inc[k]);
                        } else {
                            old[k] = a.l2_hints ? atomic_add_f64_hint(rp, inc[k], pol_keep) : atomicAdd(rp, inc[k]);
                            if (dcode[k] != dmax) du[k] = (int32_t)dcode[k];
                            else du[k] = a.l2_hints ? ld_s32_hint(&a.deg[u[k]], pol_keep) : __ldg(&a.deg[u[k]]);
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < PUSH_UB; ++k) {
                    bool cross = false;
                    if (ok[k] && !dn[k]) {
                        const double nw = old[k] + inc[k];
                        const double thr = sm.rmax[slot[k]] * (double)du[k];
                        cross = du[k] ? (old[k] < thr && nw >= thr) : (old[k] == 0.0);
                    }
                    // the queue holds one slot at a time; a step rarely spans two (tile at a slot boundary)
                    u32 pending = __ballot_sync(FULL, cross);
                    while (pending) {
                        const int s0 = __shfl_sync(FULL, slot[k], __ffs(pending) - 1);
                        if (wq && s0 != wq_slot) {
                            push_flush_warp(a, myq, wq, wq_slot, nxt, nxt_count);
                            wq = 0;
                        }
                        wq_slot = s0;
                        const u32 m = __ballot_sync(FULL, cross && slot[k] == s0);
                        if (cross && slot[k] == s0) myq[wq + __popc(m & lanemask_lt())] = make_entry(s0, (u32)du[k], u[k]);
                        wq += __popc(m);
                        pending &= ~m;
                    }
                }
                u32 nxt_step = 0;
                if (lane == 0) nxt_step = atomicAdd(&sm.step_ctr, 1u);
                step = __shfl_sync(FULL, nxt_step, 0);
            }
            const bool more = sm.G[cnt] < hi && i_cur + cnt < nf;
            __syncthreads();
            if (!more) break;
            i_cur += cnt;
        }
    }
    g_lo = g_hi;
    if (g_lo < E) cg::this_grid().sync(); // every CTA sees the same groups
    }
    if (wq) push_flush_warp(a, myq, wq, wq_slot, nxt, nxt_count);
}

// ---- dense scan of slot s after a level in which it scattered with RED: the vertices at or above their threshold ARE the
// frontier of level + 1 (phase A zeroed the previous frontier, so nothing else can be above), and their phase A happens here:
// residue zeroed, credit logged, increment written.  CTA `rank` owns a contiguous range of the vector, handled in tiles of
// SCAN_K * PUSH_THREADS vertices: pass 1 reads residue and out-degree of the tile (eight independent loads of each in flight per
// thread) and notes the hits; a block scan of the (k, warp) counts and ONE global atomic place the tile's hits in vertex order; the
// hits are compacted into shared memory SCAN_HCAP at a time (the lists alias phase B's staging arrays, idle now); pass 2 walks that
// list with all lanes busy.  With an edge list (a.el) the same batch is then expanded: exclusive scan of the hits' out-degrees, one
// global atomic for the batch's place in the slot's list, and every thread copies column words (target, pushing entry) -- the
// next level's scatters of this slot are a sequential stream.  Whole CTA, contains CTA barriers.
template <typename OffT>
__device__ __forceinline__ void push_dense_scan(const PushArgs& a, const CsrView<OffT>& g, PushSmem<OffT>& sm, int s, u32 level, u32 rank,
                                                u32 count, u64* nxt, u32* nxt_count) {
    static_assert(sizeof(sm.wqueue) >= 2 * SCAN_HCAP * sizeof(u32), "hit list + edge offsets alias the warp queues");
    static_assert(sizeof(sm.G) + sizeof(sm.inc) >= SCAN_HCAP * sizeof(u32) && offsetof(PushSmem<OffT>, inc) == offsetof(PushSmem<OffT>, G) + sizeof(sm.G),
                  "adjacency starts alias G + inc");
    u32* h_list = reinterpret_cast<u32*>(&sm.wqueue[0][0]);   // hit vertices of the batch, in vertex order
    u32* h_eoff = h_list + SCAN_HCAP;                           // out-degrees, then their exclusive prefix
    u32* h_ptr = reinterpret_cast<u32*>(&sm.G[0]);              // adjacency starts (32-bit offsets only), ~0: dangling
    const bool gather = sizeof(OffT) == 4 && a.el != nullptr;
    const int lane = lane_id(), w = threadIdx.x >> 5;
    const u32 n = (u32)a.n;
    // rows of 32 consecutive vertices are dealt to the CTAs round-robin: the relabelling puts the vertices that are hit at every level
    // (and the hubs among them) at the front of the vector, a contiguous range per CTA left CTA 0 with most of the edges to gather
    const u32 NR = (n + WARP - 1) / WARP;
    const u32 rows_cta = rank < NR ? (NR - rank + count - 1) / count : 0u;
    auto vertex_of = [&](u32 tile_row0, int k) -> u32 { // 0xffffffff: past the end
        const u32 i = tile_row0 + (u32)k * PUSH_WARPS + (threadIdx.x >> 5);
        if (i >= rows_cta) return 0xffffffffu;
        const u32 v = (i * count + rank) * WARP + (threadIdx.x & 31);
        return v < n ? v : 0xffffffffu;
    };
    double* res = a.residue + (size_t)s * n;
    const double rm = sm.rmax[s];
    const u32 lb = sm.logbase[s] + sm.prevcnt[s]; // log position of the first entry of the slot's next frontier
    const u32 par = (level + 1) & 1;
    uint4* el = gather ? a.el + ((size_t)par * a.slots + s) * a.el_cap : nullptr;
    const u32 idmask = a.colx ? ((1u << a.deg_shift) - 1u) : 0xffffffffu;
    const int32_t* __restrict__ colp = a.colx ? a.colx : g.col;
    u32 dsum_t = 0, vcnt_t = 0;
    // (development trace, CTA 0: where the scan's time goes -- stage boundaries summed per level behind the level records)
    const bool tr = a.trace && rank == 0 && level < 1024;
    u64 ts_prev = 0;
    auto stamp = [&](int stage) {
        if (!tr) return;
        u64 t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (threadIdx.x == 0 && ts_prev) a.trace[4 * 2048 + 8 * level + stage] += t - ts_prev;
        ts_prev = t;
    };
    stamp(7);
    for (u32 tb = 0; tb < rows_cta; tb += SCAN_K * PUSH_WARPS) { // tile = 512 rows
        u32 mask = 0;
        for (int k0 = 0; k0 < SCAN_K; k0 += 8) {
            double r[8];
            int32_t d[8];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const u32 v = vertex_of(tb, k0 + kk);
                r[kk] = v != 0xffffffffu ? __ldcg(&res[v]) : 0.0; // L2, never a stale L1 line: the REDs of other SMs just landed there
                d[kk] = v != 0xffffffffu ? __ldg(&a.deg[v]) : 1;  // unconditionally: one round trip instead of two
            }
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const bool hit = r[kk] > 0.0 && (d[kk] ? (r[kk] >= rm * (double)d[kk]) : true);
                const u32 bal = __ballot_sync(FULL, hit);
                if (hit) mask |= 1u << (k0 + kk);
                if (lane == 0) sm.sc_cnt[(k0 + kk) * PUSH_WARPS + w] = __popc(bal);
            }
        }
        __syncthreads();
        stamp(0); // pass 1
        // exclusive scan of the 512 (k, warp) counts in vertex order: thread t owns count t
        const u32 c = sm.sc_cnt[threadIdx.x];
        const u32 incl = warp_incl_scan(c);
        if (lane == 31) sm.sc_wtot[w] = incl;
        __syncthreads();
        const u32 wt = lane < PUSH_WARPS ? sm.sc_wtot[lane] : 0u;
        const u32 wi = warp_incl_scan(wt);
        const u32 total = __shfl_sync(FULL, wi, PUSH_WARPS - 1);
        const u32 before = w ? __shfl_sync(FULL, wi, w - 1) : 0u;
        if (threadIdx.x == 0) sm.sc_base = total ? atomicAdd(&nxt_count[s], total) : 0u;
        sm.sc_cnt[threadIdx.x] = before + incl - c; // every thread rewrites only the count it read itself
        __syncthreads();
        stamp(1); // count scan + place in the frontier
        const u32 base = sm.sc_base;
        u64* seg = nxt + (size_t)s * n;
        double* incs = a.inc + (size_t)s * n;
        for (u32 hb = 0; hb < total; hb += SCAN_HCAP) { // block-uniform; one batch unless a tile is very dense
            const u32 nb = min(SCAN_HCAP, total - hb);
#pragma unroll 8
            for (int k = 0; k < SCAN_K; ++k) { // compact the batch's hit vertices, in vertex order
                const bool hit = (mask >> k) & 1u;
                const u32 bal = __ballot_sync(FULL, hit);
                if (hit) {
                    const u32 rk = sm.sc_cnt[k * PUSH_WARPS + w] + __popc(bal & lanemask_lt()) - hb; // wraps below the batch
                    if (rk < nb) h_list[rk] = vertex_of(tb, k);
                }
            }
            __syncthreads();
            stamp(2); // compaction
            for (u32 h0 = 0; h0 < nb; h0 += 4 * PUSH_THREADS) {
                u32 v[4], d[4], pb[4];
                double r[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const u32 h = h0 + q * PUSH_THREADS + threadIdx.x;
                    v[q] = h < nb ? h_list[h] : 0xffffffffu;
                    r[q] = 0.0; d[q] = 0; pb[q] = 0;
                    if (v[q] != 0xffffffffu) {
                        r[q] = __ldcg(&res[v[q]]);
                        d[q] = (u32)__ldg(&a.deg[v[q]]);
                        if (gather) pb[q] = (u32)g.ptr[v[q]];
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (v[q] == 0xffffffffu) continue;
                    const u32 h = h0 + q * PUSH_THREADS + threadIdx.x;
                    const u32 j = base + hb + h;
                    res[v[q]] = 0.0;
                    seg[j] = make_entry(s, d[q], (int32_t)v[q]);
                    const u32 lp = lb + j;
                    if (a.log_v && lp < a.log_cap) {
                        const size_t li = (size_t)s * a.log_cap + lp;
                        a.log_v[li] = (int32_t)v[q];
                        a.log_r[li] = r[q];
                    } else {
                        double* rp = &a.reserve[(size_t)s * n + v[q]];
                        *rp = __ldcg(rp) + r[q] * a.alpha;
                    }
                    incs[j] = d[q] ? ((1.0 - a.alpha) * r[q]) / (double)d[q] : r[q] * (1.0 - a.alpha);
                    if (gather) {
                        h_eoff[h] = d[q] ? d[q] : 1u;          // a dangling vertex owns one pseudo-edge back to the source
                        h_ptr[h] = d[q] ? pb[q] : 0xffffffffu;
                    }
                    dsum_t += d[q];
                    ++vcnt_t;
                }
            }
            if (gather) {
                __syncthreads();
                stamp(3); // pass 2
                // exclusive scan of the batch's out-degrees: eight consecutive hits per thread
                u32 loc[8], tsum = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const u32 h = threadIdx.x * 8 + q;
                    loc[q] = tsum;
                    tsum += h < nb ? h_eoff[h] : 0u;
                }
                const u32 ti = warp_incl_scan(tsum);
                if (lane == 31) sm.sc_wtot[w] = ti;
                __syncthreads();
                const u32 wt2 = lane < PUSH_WARPS ? sm.sc_wtot[lane] : 0u;
                const u32 wi2 = warp_incl_scan(wt2);
                const u32 etot = __shfl_sync(FULL, wi2, PUSH_WARPS - 1);
                const u32 tbase = (w ? __shfl_sync(FULL, wi2, w - 1) : 0u) + ti - tsum;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const u32 h = threadIdx.x * 8 + q;
                    if (h < nb) h_eoff[h] = tbase + loc[q];
                }
                if (threadIdx.x == 0) {
                    const u32 eb = atomicAdd(&a.el_count[par * MAX_SLOTS + s], etot);
                    sm.sc_ebase = eb;
                    sm.sc_etot = etot;
                    if ((u64)eb + etot > a.el_cap) { a.el_bad[par * MAX_SLOTS + s] = 1; sm.sc_etot = 0; } // the level falls back to the tiles
                }
                __syncthreads();
                stamp(4); // degree scan + place in the edge list
                const u32 ebase = sm.sc_ebase, T = sm.sc_etot;
                for (u32 x0 = 0; x0 < T; x0 += 4 * PUSH_THREADS) {
                    u32 uq[4];
                    double iq[4];
                    bool okq[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const u32 x = x0 + q * PUSH_THREADS + threadIdx.x;
                        okq[q] = x < T;
                        uq[q] = 0; iq[q] = 0.0;
                        if (okq[q]) {
                            u32 l = 0, h = nb; // largest l with h_eoff[l] <= x
                            while (h - l > 1) {
                                const u32 mid = (l + h) >> 1;
                                if (h_eoff[mid] <= x) l = mid;
                                else h = mid;
                            }
                            const u32 p0 = h_ptr[l];
                            iq[q] = __ldcg(&incs[base + hb + l]); // written in pass 2 of this batch: an L2 hit, next to the column word
                            uq[q] = p0 == 0xffffffffu ? (u32)sm.source[s] : ((u32)__ldcs(&colp[p0 + (x - h_eoff[l])]) & idmask);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const u32 x = x0 + q * PUSH_THREADS + threadIdx.x;
                        if (okq[q]) {
                            const u64 ib = (u64)__double_as_longlong(iq[q]);
                            el[ebase + x] = make_uint4(uq[q], 0u, (u32)ib, (u32)(ib >> 32));
                        }
                    }
                }
            }
            __syncthreads(); // the lists are rewritten by the next batch / tile
            stamp(5); // expansion (or pass 2 without edge lists)
        }
        __syncthreads(); // sc_cnt / sc_base
    }
    const u32 ds = warp_sum(dsum_t), vc = warp_sum(vcnt_t);
    if (lane == 0 && vc) {
        atomicAdd(&sm.cnt_edges[s], ds);
        atomicAdd(&sm.cnt_verts[s], vc);
    }
}

// ---- the scatters of a slot whose edges are in its edge list: a sequential stream of (target, increment) pairs, eight of them in flight
// per thread, fp64 RED (the scan that follows finds the next frontier).  Adds to the EL_HOT hottest vertices (lowest ids after the
// relabelling: ~15 % of all adds, and the few 128-byte lines on which the L2 would serialise them) are collected in shared memory --
// the accumulators alias the warp queues, idle now -- and leave as one RED per CTA and vertex.  Contiguous share per CTA.
constexpr u32 EL_HOT = 4096;
template <typename OffT>
__device__ __forceinline__ void push_el_adds(const PushArgs& a, PushSmem<OffT>& sm, int s, u32 level, u32 rank, u32 count) {
    static_assert(sizeof(sm.wqueue) >= EL_HOT * sizeof(double), "hot accumulators alias the warp queues");
    double* acc = reinterpret_cast<double*>(&sm.wqueue[0][0]);
    const u32 par = level & 1;
    const u32 E = min(*(volatile u32*)&a.el_count[par * MAX_SLOTS + s], a.el_cap);
    const uint4* __restrict__ el = a.el + ((size_t)par * a.slots + s) * a.el_cap;
    double* res = a.residue + (size_t)s * a.n;
    const u32 lo = (u32)(((u64)E * rank) / count), hi = (u32)(((u64)E * (rank + 1)) / count);
    const u64 pol_keep = l2_policy_evict_last();
    const u32 hot = min(EL_HOT, (u32)a.n);
    // the hot front of the vector first: a hot line that has to be fetched from DRAM at its first touch stalls its L2 slice for the
    // whole fill while hundreds of adds queue behind it (cold vector + graph targets: 159 us per 3.4 M adds, cold + uniform targets 46 us)
    {
        const u32 lines = min(a.el_prefetch, (u32)a.n) / 16u;
        for (u32 i = rank * PUSH_THREADS + threadIdx.x; i < lines; i += count * PUSH_THREADS)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(res + (size_t)i * 16u));
    }
    for (u32 i = threadIdx.x; i < hot; i += PUSH_THREADS) acc[i] = 0.0;
    __syncthreads();
    double dummy = 0.0;
    for (u32 x0 = lo + threadIdx.x; x0 < hi; x0 += 8 * PUSH_THREADS) {
        uint4 e[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            u32 x = x0 + q * PUSH_THREADS;
            if ((a.debug_el & 8u) && x < hi) x = lo + (u32)(((u64)(x - lo) * 7919u) % (u64)(hi - lo)); // ablation: entries in a scattered order
            e[q] = x < hi ? __ldcs(&el[x]) : make_uint4(0xffffffffu, 0u, 0u, 0u);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (e[q].x != 0xffffffffu && !(a.debug_el & 2u) && e[q].x >= a.debug_el_skip) {
                const double inc = __longlong_as_double((long long)(((u64)e[q].w << 32) | e[q].z));
                if (a.debug_el & 1u) e[q].x = (u32)(((u64)((x0 + q * PUSH_THREADS) * 2654435761u) * (u32)a.n) >> 32);
                if (a.debug_el & 64u) { // ablation: random targets with the generator's in-degree law, rank == id (scripts/ubench_hot.cu)
                    const u32 h1 = (x0 + q * PUSH_THREADS) * 2654435761u + 12345u, h2 = (h1 ^ (h1 >> 15)) * 0x846ca68bu;
                    const double uu = ((double)(h2 ^ (h2 >> 16)) + 0.5) * (1.0 / 4294967296.0);
                    const double e1 = 1.0 - 1.0 / 1.3, aa = pow(50.0, e1), bb = pow((double)a.n + 50.0, e1);
                    const double xx = pow(uu * (bb - aa) + aa, 1.0 / e1) - 50.0;
                    e[q].x = xx < 0 ? 0u : min((u32)xx, (u32)a.n - 1u);
                }
                if (a.debug_el & 32u) e[q].x = (u32)(((u64)e[q].x * 1000003ull) % (u64)(u32)a.n); // ablation: same heat per vertex, hot vertices spread over the lines
                if (e[q].x < hot && !(a.debug_el & 4u)) atomicAdd(&acc[e[q].x], inc);
                else if (a.debug_el & 16u) dummy += atomicAdd(&res[e[q].x], inc); // ablation: ATOM with return instead of RED
                else if (a.l2_hints) red_add_f64_hint(&res[e[q].x], inc, pol_keep);
                else atomicAdd(&res[e[q].x], inc);
            }
    }
    if (dummy == 123.456e300) acc[0] = dummy;
    __syncthreads();
    for (u32 i = threadIdx.x; i < hot; i += PUSH_THREADS) {
        const double v = acc[i];
        if (v != 0.0) atomicAdd(&res[i], v);
    }
    __syncthreads(); // the queues' memory is handed back
}

template <typename OffT, bool DN>
__global__ void __launch_bounds__(PUSH_THREADS, 2) push_kernel(PushArgs a, CsrView<OffT> g) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char push_smem_raw[];
    PushSmem<OffT>& sm = *reinterpret_cast<PushSmem<OffT>*>(push_smem_raw);
    PushCtl* ctl = a.ctl;
    for (int i = threadIdx.x; i < MAX_SLOTS; i += blockDim.x) {
        sm.rmax[i] = i < a.slots ? a.rmax[i] : 0.0;
        sm.source[i] = i < a.slots ? a.source[i] : 0;
        sm.cnt_edges[i] = 0;
        sm.cnt_verts[i] = 0;
        sm.logbase[i] = (a.log_v && i < a.slots) ? a.log_cur[i] : 0u;
        sm.prevcnt[i] = 0;
        sm.dense[i] = 0;
        sm.adone[i] = 0;
    }
    __syncthreads();

    for (u32 level = 0;; ++level) {
        // the level's frontier: per-slot segments concatenated in slot order
        if (threadIdx.x < WARP) { // exclusive scan of the (<= 64) per-slot counts by one warp
            const int l = threadIdx.x;
            const u32 c0 = l < a.slots ? *((volatile u32*)&ctl->fcount[level % 3][l]) : 0u;
            const u32 c1 = l + WARP < a.slots ? *((volatile u32*)&ctl->fcount[level % 3][l + WARP]) : 0u;
            // a slot that scattered with RED was scanned: its frontier is ready-made; large frontiers scatter with RED this time; a slot
            // that was scanned AND is large again has its edges in its edge list and takes no part in phase A / phase B
            const bool d0 = DN && c0 >= a.dense_min && c0 > 0, d1 = DN && c1 >= a.dense_min && c1 > 0;
            const bool a0 = l < a.slots && sm.dense[l], a1 = l + WARP < a.slots && sm.dense[l + WARP]; // dense at the previous level
            bool e0 = false, e1 = false;
            if (DN && a.el) {
                const u32 par = level & 1;
                if (a0 && d0) e0 = *((volatile u32*)&a.el_bad[par * MAX_SLOTS + l]) == 0;
                if (a1 && d1) e1 = *((volatile u32*)&a.el_bad[par * MAX_SLOTS + l + WARP]) == 0;
            }
            const u32 f0 = e0 ? 0u : c0, f1 = e1 ? 0u : c1; // entries that go through phase A / phase B
            const u32 i0 = warp_incl_scan(f0);
            const u32 i1 = warp_incl_scan(f1) + __shfl_sync(FULL, i0, 31);
            if (l < a.slots) { sm.fbase[l] = i0 - f0; sm.logbase[l] += sm.prevcnt[l]; sm.prevcnt[l] = c0; }
            if (l + WARP < a.slots) { sm.fbase[l + WARP] = i1 - f1; sm.logbase[l + WARP] += sm.prevcnt[l + WARP]; sm.prevcnt[l + WARP] = c1; }
            if (l == 31) sm.fbase[a.slots] = i1;
            if (l < a.slots) { sm.adone[l] = a0; sm.dense[l] = d0; sm.el_ok[l] = e0; }
            if (l + WARP < a.slots) { sm.adone[l + WARP] = a1; sm.dense[l + WARP] = d1; sm.el_ok[l + WARP] = e1; }
            const u32 nd = __popc(__ballot_sync(FULL, d0)) + __popc(__ballot_sync(FULL, d1));
            const u32 anyc = __ballot_sync(FULL, c0 != 0 || c1 != 0);
            if (l == 0) { sm.ndense = nd; sm.nf_all = anyc ? 1u : 0u; }
            if (blockIdx.x == 0) { // levels in which the slot pushed something
                const int lvl = (int)(a.level_base + level + 1);
                if (level < a.max_levels) {
                    if (l < a.slots && c0 && a.lastlvl[l] < lvl) { a.lastlvl[l] = lvl; atomicAdd(&a.levels[l], 1ull); }
                    if (l + WARP < a.slots && c1 && a.lastlvl[l + WARP] < lvl) { a.lastlvl[l + WARP] = lvl; atomicAdd(&a.levels[l + WARP], 1ull); }
                }
            }
        }
        __syncthreads();
        const u32 nf = sm.fbase[a.slots]; // (may be 0 while slots with an edge list still have work)
        const u32 ndense = sm.ndense;     // (read now: warp 0 writes the next level's value while slower warps are still behind the last barrier)
        if (!sm.nf_all) break;
        if (level >= a.max_levels) { // never reached in practice (2^20 levels): report instead of dropping the frontier silently
            if (blockIdx.x == 0 && threadIdx.x == 0 && a.err) *a.err = 1;
            if (threadIdx.x < (u32)a.slots) sm.prevcnt[threadIdx.x] = 0; // the counts just read belong to a level that did not run: not logged
            __syncthreads();
            break;
        }
        const u64* cur = (level & 1) ? a.front1 : a.front0; // written by the previous level: L2 reads only
        u64* nxt = (level & 1) ? a.front0 : a.front1;
        u32* nxt_count = ctl->fcount[(level + 1) % 3];
        if (blockIdx.x == 0) {
            if (threadIdx.x < (u32)a.slots) {
                ctl->fcount[(level + 2) % 3][threadIdx.x] = 0;
                if (DN && a.el) { // the lists of the next level's parity were consumed a level ago
                    a.el_count[((level + 1) & 1) * MAX_SLOTS + threadIdx.x] = 0;
                    a.el_bad[((level + 1) & 1) * MAX_SLOTS + threadIdx.x] = 0;
                }
            }
            if (threadIdx.x == 0) {
                ctl->levels_run = level + 1;
                if (a.trace && level < a.trace_cap) {
                    u64 t;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                    a.trace[4 * level] = t;
                    a.trace[4 * level + 1] = nf | ((u64)sm.ndense << 40);
                }
            }
        }
        push_phase_a<OffT, DN>(a, sm, cur, nf, level, blockIdx.x, gridDim.x);
        grid.sync();
        if (blockIdx.x == 0 && threadIdx.x == 0 && a.trace && level < a.trace_cap) {
            u64 t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            a.trace[4 * level + 3] = t;
        }
        push_phase_b<OffT, DN>(a, g, sm, cur, nf, level, blockIdx.x, gridDim.x, nxt, nxt_count);
        grid.sync();
        if (DN && ndense) {
            // the dense slots, one after the other with the whole grid (their vectors stay in the L2 from the adds to the scan): a slot
            // with an edge list streams its scatters now; then one pass over the residue vector finds the next frontier, does its
            // phase A and lists its edges
            for (int s = 0; s < a.slots; ++s) {
                if (!sm.dense[s]) continue;
                // (development trace, CTA 0: own adds / wait / own scan + gather / wait, summed over the level's dense slots)
                const bool tr = a.trace && blockIdx.x == 0 && level < 1024;
                u64 t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
                if (tr) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                if (sm.el_ok[s]) {
                    push_el_adds<OffT>(a, sm, s, level, blockIdx.x, gridDim.x);
                    if (tr) { __syncthreads(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); }
                    grid.sync();
                } else if (tr) t1 = t0;
                if (tr) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2));
                push_dense_scan<OffT>(a, g, sm, s, level, blockIdx.x, gridDim.x, nxt, nxt_count);
                if (tr) { __syncthreads(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t3)); }
                if (a.el) grid.sync(); // (without edge lists the scans of different slots touch different vectors: one barrier at the end)
                if (tr) {
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t4));
                    if (threadIdx.x == 0) {
                        u64* tx = a.trace + 4 * 1024 + 4 * level;
                        tx[0] += t1 - t0; tx[1] += t2 - t1; tx[2] += t3 - t2; tx[3] += t4 - t3;
                    }
                }
            }
            if (!a.el) grid.sync();
            __syncthreads(); // nobody is still reading this level's flags when warp 0 writes the next level's
        }
    }
    __syncthreads();
    push_flush_counters(a, sm); // what the last scans counted
    // every CTA holds the same log positions; CTA 0 publishes them for the next launch of the wave / the apply pass
    if (a.log_v && blockIdx.x == 0 && threadIdx.x < (u32)a.slots) a.log_cur[threadIdx.x] = sm.logbase[threadIdx.x] + sm.prevcnt[threadIdx.x];
}

// reserve[v] += alpha * r for every logged push; grid.y = slot, so the grid works through the slots one after another
// and each slot's 8n-byte reserve vector is L2-resident while its REDs land.
__global__ void __launch_bounds__(256) apply_log_kernel(int32_t n, double alpha, const u32* __restrict__ log_cur, u32 cap,
                                                         const int32_t* __restrict__ log_v, const double* __restrict__ log_r,
                                                         double* __restrict__ reserve) {
    const int slot = blockIdx.y;
    const u32 cnt = min(log_cur[slot], cap);
    const int32_t* __restrict__ lv = log_v + (size_t)slot * cap;
    const double* __restrict__ lr = log_r + (size_t)slot * cap;
    double* res = reserve + (size_t)slot * n;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) atomicAdd(&res[__ldcs(&lv[i])], __ldcs(&lr[i]) * alpha);
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
        if (seed_source) {
            front0[(size_t)slot * n] = make_entry(slot, (u32)deg[s], s);
            ctl->fcount[0][slot] = 1;
        }
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
        int32_t d = 0;
        if (v < n) {
            const double r = res[v];
            if (r > 0.0) {
                d = deg[v];
                pred = d ? (r >= rm * (double)d) : true;
            }
        }
        warp_append<u64>(pred, make_entry(slot, (u32)d, v), front0 + (size_t)slot * n, &ctl->fcount[0][slot]);
    }
}

// ---------------------------------------------------------------------------------------------
// Deterministic per-slot reduction of the residue vector: rsum (fixed summation tree) and the
// number of non-zero entries.  grid = (blocks, slots).
// ---------------------------------------------------------------------------------------------
constexpr int RED_THREADS = 256;
// The same dense pass also prepares the NEXT resumable round speculatively: every vertex with
// residue >= next_rmax*d_out (dangling: any positive residue) is appended to the slot's seed segment, so a
// round that does continue starts without a second scan of the residue vectors.  A block counts its seeds first and
// reserves their places with ONE global atomic (a per-warp atomic on the per-slot counter serialises at the L2 and
// made this pass 7x slower than the plain reduction), then writes them in a second sweep over its (L2-hot) range.
__global__ void __launch_bounds__(RED_THREADS) residue_partial_kernel(int32_t n, const double* __restrict__ residue,
                                                                      double* __restrict__ part_sum,
                                                                      u32* __restrict__ part_nnz,
                                                                      const int32_t* __restrict__ deg,
                                                                      const double* __restrict__ next_rmax, // [slots], <= 0: no seeding
                                                                      u64* __restrict__ seeds, u32* __restrict__ seed_count) {
    __shared__ double s_sum[RED_THREADS];
    __shared__ u32 s_nnz[RED_THREADS];
    __shared__ u32 s_seed[RED_THREADS];
    __shared__ u32 s_base;
    const int slot = blockIdx.y;
    const double* __restrict__ res = residue + (size_t)slot * n;
    const int per_block = (n + gridDim.x - 1) / gridDim.x;
    const int lo = min(n, blockIdx.x * per_block), hi = min(n, lo + per_block);
    const double rm = next_rmax ? next_rmax[slot] : 0.0;
    double s = 0.0;
    u32 c = 0, ns = 0;
    const bool seeding = rm > 0.0;
    auto take = [&](double r, int32_t d) { // same per-thread order as a plain strided loop: rsum stays bit-identical
        s += r;
        c += r > 0.0;
        if (seeding && r > 0.0) ns += d ? (r >= rm * (double)d) : 1u;
    };
    int v = lo + threadIdx.x;
    for (; v + 3 * RED_THREADS < hi; v += 4 * RED_THREADS) { // four independent loads in flight per thread
        const double r0 = res[v], r1 = res[v + RED_THREADS], r2 = res[v + 2 * RED_THREADS],
                     r3 = res[v + 3 * RED_THREADS];
        // the out-degrees next to them, unconditionally when seeds are wanted (a load behind `r > 0` made every element two
        // dependent round trips; the degree vector is shared by all slots and stays in the L2)
        int32_t d0 = 0, d1 = 0, d2 = 0, d3 = 0;
        if (seeding) { d0 = deg[v]; d1 = deg[v + RED_THREADS]; d2 = deg[v + 2 * RED_THREADS]; d3 = deg[v + 3 * RED_THREADS]; }
        take(r0, d0);
        take(r1, d1);
        take(r2, d2);
        take(r3, d3);
    }
    for (; v < hi; v += RED_THREADS) take(res[v], seeding ? deg[v] : 0);
    s_sum[threadIdx.x] = s;
    s_nnz[threadIdx.x] = c;
    s_seed[threadIdx.x] = ns;
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
    if (rm <= 0.0) return; // uniform per block
    // exclusive scan of the per-thread seed counts (thread t owns elements lo+t, lo+t+256, ...)
    if (threadIdx.x == 0) {
        u32 acc = 0;
        for (int t = 0; t < RED_THREADS; ++t) { const u32 x = s_seed[t]; s_seed[t] = acc; acc += x; }
        s_base = acc ? atomicAdd(&seed_count[slot], acc) : 0u;
    }
    __syncthreads();
    u64* out = seeds + (size_t)slot * n + s_base + s_seed[threadIdx.x];
    int v2 = lo + threadIdx.x;
    auto emit = [&](double r, int32_t d, int vv) {
        if (r > 0.0 && (d ? (r >= rm * (double)d) : true)) *out++ = make_entry(slot, (u32)d, vv);
    };
    for (; v2 + 3 * RED_THREADS < hi; v2 += 4 * RED_THREADS) { // same order as the count above, four elements in flight
        const double r0 = res[v2], r1 = res[v2 + RED_THREADS], r2 = res[v2 + 2 * RED_THREADS], r3 = res[v2 + 3 * RED_THREADS];
        const int32_t d0 = deg[v2], d1 = deg[v2 + RED_THREADS], d2 = deg[v2 + 2 * RED_THREADS], d3 = deg[v2 + 3 * RED_THREADS];
        emit(r0, d0, v2);
        emit(r1, d1, v2 + RED_THREADS);
        emit(r2, d2, v2 + 2 * RED_THREADS);
        emit(r3, d3, v2 + 3 * RED_THREADS);
    }
    for (; v2 < hi; v2 += RED_THREADS) emit(res[v2], deg[v2], v2);
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
