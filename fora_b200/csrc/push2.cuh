// fora_b200/csrc/push2.cuh -- forward push, second generation: L2-resident sub-waves + per-slot tails (sm_100a).
//
// Same per-vertex rule and the same level-synchronous schedule as push.cuh (/root/reference/algo.h:954-1093; the parity tests
// compare against a CPU restatement of exactly this schedule), executed differently:
//
//   * A wave of S query slots is pushed in SUB-WAVES of k slots (k = 2..3 at LiveJournal scale): k dense fp64 residue vectors
//     (39 MB each) stay resident in the 126 MB L2 for all levels of the sub-wave, so every scatter is an L2 atomic (measured
//     ceiling 125 G/s up to ~100 MB footprint, profiles/r2_ubench_push.txt) instead of a DRAM sector round trip (24 G/s).
//   * push2_kernel (cooperative, persistent over levels) runs the levels whose frontier is large.  Phase B is WARP-AUTONOMOUS:
//     a warp takes 32 frontier entries, scans their out-degrees in registers, and walks the concatenated edge range with
//     P2_UB coalesced column loads and atomics in flight per lane -- no CTA barrier, no global scan, no tile search, so warps
//     drift apart and one warp's column-load wait overlaps another's atomics (the first-generation kernel staged batches behind
//     CTA barriers: 17 % issue-active with everything L2-resident, profiles/r2a_ncu_push_s2_lines.txt).  Vertices with
//     >= P2_HUB_DEG out-edges are listed in phase A and cut across the whole grid.
//   * push_tail_kernel runs the levels whose frontier is small (the head and the long tail of every round: half of all levels)
//     with ONE CTA per slot and CTA-local barriers, all slots of the wave concurrently: ~3 us per level instead of two grid
//     barriers plus two dependent phases of a 148-CTA grid.
//   The host alternates tail -> sub-waves -> tail until every frontier is empty (engine.cu launch_push2).
//
// Frontier hand-over between kernels: slot s keeps its current frontier in buffer ctl->par[s] (front0 / front1 segment of the
// slot) with ctl->fcount[par][s] entries, and ctl->fcount[par^1][s] == 0.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "push.cuh"

namespace fora {

constexpr int P2_THREADS = 1024;
constexpr int P2_WARPS = P2_THREADS / WARP;
#ifndef CFG_P2_UB
#define CFG_P2_UB 4
#endif
constexpr int P2_UB = CFG_P2_UB;       // edges in flight per lane
#ifndef CFG_P2_UA
#define CFG_P2_UA 4
#endif
constexpr int P2_UA = CFG_P2_UA;       // frontier entries in flight per thread in phase A
constexpr int P2_WQ = 256;             // per-warp queue of crossing vertices
#ifndef CFG_P2_HUB_DEG
#define CFG_P2_HUB_DEG 256
#endif
constexpr u32 P2_HUB_DEG = CFG_P2_HUB_DEG; // out-degree from which a frontier vertex is cut across the whole grid (a warp walks <= 2 steps per entry)
constexpr u32 P2_HUB_CAP = 1024;       // hub entries per level = what one CTA stages in shared memory, one per thread (a full list leaves
                                       // the rest to the per-warp path: slow, correct)
constexpr u32 P2_HUB_SM = P2_HUB_CAP;
constexpr u32 P2_CYC = 4;              // phase B hands 32-entry groups to the CTAs block-cyclically, P2_CYC consecutive groups at a time
constexpr int P2_MAX_SUB = 8;          // slots per sub-wave handled by one push2 launch
constexpr int TAIL_THREADS = 1024;
constexpr u32 TAIL_NF_CAP = 2048;      // frontier entries a tail CTA keeps in shared memory

struct Push2Args {
    PushArgs p;            // shared with the first-generation kernel (vectors, frontier buffers, counters, credit log, packed columns)
    int32_t slot0, k;      // this launch handles slots [slot0, slot0 + k)
    u32 tail_nf, tail_e;   // a level with <= tail_nf entries and <= tail_e edges belongs to the tail kernel
    double* rv;            // per frontier entry (global index within the launch): residue pushed; < 0: listed as a hub
    void* begs;            // OffT per frontier entry: start of the vertex's adjacency list (loaded in phase A, off phase B's critical path)
    u32* hub;              // [2 * P2_HUB_CAP] frontier indices of hub entries, by level parity
    u32* hub_cnt;          // [2]
    u32* force;            // [MAX_SLOTS] set by the tail kernel when it refused a level for its edge count
    u32* left;             // [MAX_SLOTS] tail kernel: frontier entries left when it returned (0: the slot's round is complete)
    int32_t* err;          // [1] set on an internal limit (level cap)
    int prefetch;          // 1: stream the sub-wave's residue vectors into the L2 before the first level
};

struct P2Smem {
    u64 wqueue[P2_WARPS][P2_WQ];
    // per-warp staging of the 32 entries a warp is expanding
    u32 w_off[P2_WARPS][WARP + 1];
    u32 w_beg32[P2_WARPS][WARP];   // low / high halves of the adjacency start (OffT may be 64-bit)
    u32 w_beghi[P2_WARPS][WARP];
    double w_inc[P2_WARPS][WARP];
    int w_slot[P2_WARPS][WARP];    // slot (relative to slot0), or ~slot when the entry is dangling
    double rmax[P2_MAX_SUB];
    int32_t source[P2_MAX_SUB];
    u32 par[P2_MAX_SUB];
    u32 cnt[P2_MAX_SUB];
    u32 fbase[P2_MAX_SUB + 1];
    u32 logpos[P2_MAX_SUB];
    u32 acc_edges[P2_MAX_SUB], acc_verts[P2_MAX_SUB]; // this CTA's share, flushed once at kernel exit
    u32 next_group;
    int anybig;
    // hub phase: up to P2_HUB_SM listed entries, their edges laid on one line (hb_off = exclusive prefix of the out-degrees)
    u32 hb_off[P2_HUB_SM + 1];
    u32 hb_beg32[P2_HUB_SM], hb_beghi[P2_HUB_SM];
    double hb_inc[P2_HUB_SM];
    int hb_slot[P2_HUB_SM];
    u32 hb_wtot[P2_WARPS];
};

__device__ __forceinline__ void st_stream_f64(double* p, double v) { __stcs(p, v); }

__device__ __forceinline__ void prefetch_l2_keep(const void* p) {
    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p));
}

// slot (relative) and segment offset of global frontier index i
__device__ __forceinline__ int p2_slot_of(const P2Smem& sm, int k, u32 i) {
    int s = 0;
#pragma unroll
    for (int t = 1; t < P2_MAX_SUB; ++t)
        if (t < k && sm.fbase[t] <= i) s = t;
    return s;
}
__device__ __forceinline__ const u64* p2_front(const PushArgs& p, int abs_slot, u32 buf) {
    return (buf ? p.front1 : p.front0) + (size_t)abs_slot * p.n;
}

// append this warp's queue (entries of ONE relative slot) to that slot's next frontier
__device__ __forceinline__ void p2_flush(const Push2Args& a, const P2Smem& sm, PushCtl* ctl, const u64* myq, u32 wq, int rs) {
    const int lane = lane_id();
    const int as = a.slot0 + rs;
    const u32 nb = sm.par[rs] ^ 1u;
    u32 base = 0;
    __syncwarp();
    if (lane == 0) base = atomicAdd(&ctl->fcount[nb][as], wq);
    base = __shfl_sync(FULL, base, 0);
    u64* dst = const_cast<u64*>(p2_front(a.p, as, nb)) + base;
    for (u32 t = lane; t < wq; t += WARP) dst[t] = myq[t];
    __syncwarp();
}

// One warp scatters the edges [x_begin, x_end) of an edge line described by staging arrays in shared memory: owner o holds
// [off[o], off[o+1]) (offsets ascending, off[nown] = total).  WIDE = false: the warp's own 32 entries (5-step search);
// WIDE = true: the CTA-wide hub arrays (nown <= P2_HUB_SM).  P2_UB column loads, then P2_UB atomics, in flight per lane.
template <typename OffT, bool WIDE>
__device__ __forceinline__ void p2_scatter(const Push2Args& a, const CsrView<OffT>& g, P2Smem& sm, PushCtl* ctl, int w, u32 x_begin, u32 x_end,
                                           u32& wq, int& wq_slot, u64 pol_keep, u64 pol_stream, u32 nown = WARP) {
    const PushArgs& p = a.p;
    const int lane = lane_id();
    u64* myq = sm.wqueue[w];
    const u32* off = WIDE ? sm.hb_off : sm.w_off[w];
    const u32* beg32 = WIDE ? sm.hb_beg32 : sm.w_beg32[w];
    const u32* beghi = WIDE ? sm.hb_beghi : sm.w_beghi[w];
    const double* incs = WIDE ? sm.hb_inc : sm.w_inc[w];
    const int* slots = WIDE ? sm.hb_slot : sm.w_slot[w];
    const u32 dmax = p.colx ? (0xffffffffu >> p.deg_shift) : 0u, idmask = p.colx ? ((1u << p.deg_shift) - 1u) : 0xffffffffu;
    const int32_t* __restrict__ colp = p.colx ? p.colx : g.col;
    for (u32 x0 = x_begin; x0 < x_end; x0 += WARP * P2_UB) {
        if (wq > P2_WQ - WARP * P2_UB) {
            p2_flush(a, sm, ctl, myq, wq, wq_slot);
            wq = 0;
        }
        int rs[P2_UB];
        int32_t u[P2_UB];
        u32 dcode[P2_UB];
        double inc[P2_UB], old[P2_UB];
        bool ok[P2_UB];
#pragma unroll
        for (int k = 0; k < P2_UB; ++k) {
            const u32 x = x0 + (u32)k * WARP + lane;
            ok[k] = x < x_end;
            rs[k] = 0; u[k] = 0; inc[k] = 0.0; dcode[k] = dmax;
            if (ok[k]) {
                u32 o = 0; // largest owner with off[o] <= x
                if (WIDE) {
                    u32 hi = nown;
                    while (hi - o > 1) {
                        const u32 mid = (o + hi) >> 1;
                        if (off[mid] <= x) o = mid;
                        else hi = mid;
                    }
                } else {
                    if (off[16] <= x) o = 16;
                    if (off[o + 8] <= x) o += 8;
                    if (off[o + 4] <= x) o += 4;
                    if (off[o + 2] <= x) o += 2;
                    if (off[o + 1] <= x) o += 1;
                }
                const int sj = slots[o];
                rs[k] = sj < 0 ? ~sj : sj;
                inc[k] = incs[o];
                if (sj < 0) {
                    u[k] = sm.source[rs[k]];
                } else {
                    const OffT beg = (OffT)(((u64)beghi[o] << 32) | beg32[o]);
                    const int32_t* cp = &colp[beg + (OffT)(x - off[o])];
                    const u32 raw = (u32)(p.l2_hints ? ld_col_stream(cp, pol_stream) : __ldcs(cp));
                    u[k] = (int32_t)(raw & idmask);
                    if (p.colx) dcode[k] = raw >> p.deg_shift;
                }
            }
        }
        int32_t du[P2_UB];
#pragma unroll
        for (int k = 0; k < P2_UB; ++k) {
            du[k] = 0;
            if (ok[k]) {
                double* rp = &p.residue[(size_t)(a.slot0 + rs[k]) * p.n + u[k]];
                old[k] = p.l2_hints ? atomic_add_f64_hint(rp, inc[k], pol_keep) : atomicAdd(rp, inc[k]);
                if (dcode[k] != dmax) du[k] = (int32_t)dcode[k];
                else du[k] = p.l2_hints ? ld_s32_hint(&p.deg[u[k]], pol_keep) : __ldg(&p.deg[u[k]]);
            }
        }
#pragma unroll
        for (int k = 0; k < P2_UB; ++k) {
            bool cross = false;
            if (ok[k]) {
                const double nw = old[k] + inc[k];
                const double thr = sm.rmax[rs[k]] * (double)du[k];
                cross = du[k] ? (old[k] < thr && nw >= thr) : (old[k] == 0.0);
            }
            u32 pending = __ballot_sync(FULL, cross);
            while (pending) { // the queue holds one slot at a time; a step rarely spans two
                const int s0 = __shfl_sync(FULL, rs[k], __ffs(pending) - 1);
                if (wq && s0 != wq_slot) {
                    p2_flush(a, sm, ctl, myq, wq, wq_slot);
                    wq = 0;
                }
                wq_slot = s0;
                const u32 m = __ballot_sync(FULL, cross && rs[k] == s0);
                if (cross && rs[k] == s0) myq[wq + __popc(m & lanemask_lt())] = make_entry(a.slot0 + s0, (u32)du[k], u[k]);
                wq += __popc(m);
                pending &= ~m;
            }
        }
    }
}

template <typename OffT>
__global__ void __launch_bounds__(P2_THREADS, 1) push2_kernel(Push2Args a, CsrView<OffT> g) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char p2_smem_raw[];
    P2Smem& sm = *reinterpret_cast<P2Smem*>(p2_smem_raw);
    const PushArgs& p = a.p;
    PushCtl* ctl = p.ctl;
    const int lane = lane_id(), w = threadIdx.x >> 5;
    const int k = a.k;
    const u32 G = gridDim.x, c = blockIdx.x;
    if (threadIdx.x < P2_MAX_SUB) {
        const int t = threadIdx.x, as = a.slot0 + t;
        const bool in = t < k;
        sm.rmax[t] = in ? p.rmax[as] : 0.0;
        sm.source[t] = in ? p.source[as] : 0;
        sm.par[t] = in ? ctl->par[as] : 0u;
        sm.logpos[t] = (in && p.log_v) ? p.log_cur[as] : 0u;
        sm.acc_edges[t] = 0;
        sm.acc_verts[t] = 0;
    }
    __syncthreads();
    const u64 pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
    u32 wq = 0;
    int wq_slot = 0;
    u32 levels_run = 0;

    for (u32 level = 0;; ++level) {
        // ---- the level's frontier: the k slots' segments concatenated in slot order
        if (threadIdx.x < P2_MAX_SUB) {
            const int t = threadIdx.x;
            sm.cnt[t] = t < k ? *((volatile u32*)&ctl->fcount[sm.par[t]][a.slot0 + t]) : 0u;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            u32 acc = 0;
            int big = 0;
            for (int t = 0; t < k; ++t) {
                sm.fbase[t] = acc;
                acc += sm.cnt[t];
                if (sm.cnt[t] > a.tail_nf || (sm.cnt[t] && *((volatile u32*)&a.force[a.slot0 + t]))) big = 1;
            }
            for (int t = k; t <= P2_MAX_SUB; ++t) sm.fbase[t] = acc;
            sm.anybig = big;
            sm.next_group = 0;
        }
        __syncthreads();
        const u32 nf = sm.fbase[k];
        if (nf == 0 || !sm.anybig) break;
        if (level >= p.max_levels) {
            if (c == 0 && threadIdx.x == 0) *a.err = 1;
            break;
        }
        ++levels_run;
        if (level == 0 && a.prefetch) { // stream the sub-wave's residue vectors into the L2 (cheaper than ~1.2 M random first-touch misses per slot)
            for (int t = 0; t < k; ++t) {
                if (!sm.cnt[t]) continue;
                const char* base = (const char*)(p.residue + (size_t)(a.slot0 + t) * p.n);
                const size_t lines = ((size_t)p.n * sizeof(double) + 127) / 128;
                for (size_t l = (size_t)c * P2_THREADS + threadIdx.x; l < lines; l += (size_t)G * P2_THREADS) prefetch_l2_keep(base + l * 128);
            }
        }
        const u32 cs = (nf + G - 1) / G;
        const u32 lo_i = min(nf, c * cs), hi_i = min(nf, lo_i + cs);
        const u32 hp = level & 1u;
        if (c == 0) {
            if (threadIdx.x < (u32)k) {
                const int t = threadIdx.x, as = a.slot0 + t;
                ctl->fcount[sm.par[t] ^ 1u][as] = 0; // the buffer the previous level consumed becomes this level's output
                if (sm.cnt[t]) p.levels[as] += 1;    // only this thread of this CTA touches it
            }
            if (threadIdx.x == 0) {
                a.hub_cnt[hp ^ 1u] = 0;
                if (p.trace && level < p.trace_cap) {
                    u64 t;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                    p.trace[4 * level] = t;
                    p.trace[4 * level + 1] = nf;
                }
            }
        }
        // ---- phase A: read + zero the residues, log the reserve credits, list the hubs
        for (u32 b0 = lo_i; b0 < hi_i; b0 += P2_THREADS * P2_UA) {
            u64 e[P2_UA];
            double r[P2_UA];
            u32 idx[P2_UA];
            int rs[P2_UA];
#pragma unroll
            for (int q = 0; q < P2_UA; ++q) {
                idx[q] = b0 + (u32)q * P2_THREADS + threadIdx.x;
                e[q] = ~0ull;
                rs[q] = 0;
                if (idx[q] < hi_i) {
                    rs[q] = p2_slot_of(sm, k, idx[q]);
                    e[q] = __ldcg(p2_front(p, a.slot0 + rs[q], sm.par[rs[q]]) + (idx[q] - sm.fbase[rs[q]])); // written by other SMs last level: L2
                }
            }
#pragma unroll
            for (int q = 0; q < P2_UA; ++q) {
                r[q] = 0.0;
                if (e[q] != ~0ull)
                    r[q] = __longlong_as_double((long long)atomicExch((unsigned long long*)&p.residue[(size_t)(a.slot0 + rs[q]) * p.n + (u32)e[q]], 0ull));
            }
            u32 dsum = 0, vcnt = 0;
            int slot_first = -1;
            bool same = true;
#pragma unroll
            for (int q = 0; q < P2_UA; ++q) {
                if (e[q] == ~0ull) continue;
                const int as = a.slot0 + rs[q];
                const int32_t v = entry_vertex(e[q]);
                u32 d = entry_deg24(e[q]);
                if (d == DEG_SAT) d = (u32)__ldg(&p.deg[v]);
                const u32 lp = sm.logpos[rs[q]] + (idx[q] - sm.fbase[rs[q]]);
                if (p.log_v && lp < p.log_cap) {
                    const size_t li = (size_t)as * p.log_cap + lp;
                    __stcs(&p.log_v[li], v);
                    __stcs(&p.log_r[li], r[q]);
                } else { // no room (or no log): direct update of the reserve
                    const size_t gi = (size_t)as * p.n + v;
                    p.reserve[gi] = __ldcg(&p.reserve[gi]) + r[q] * p.alpha;
                }
                __stcs(&((OffT*)a.begs)[idx[q]], d ? g.ptr[v] : (OffT)0);
                double out = r[q];
                if (d >= P2_HUB_DEG) {
                    const u32 pos = atomicAdd(&a.hub_cnt[hp], 1u);
                    if (pos < P2_HUB_CAP) {
                        a.hub[hp * P2_HUB_CAP + pos] = idx[q];
                        out = -r[q];
                    }
                }
                st_stream_f64(&a.rv[idx[q]], out);
                dsum += d;
                ++vcnt;
                if (slot_first < 0) slot_first = rs[q];
                else same = same && rs[q] == slot_first;
            }
            // per-slot work counters (cost model of --balanced, roofline accounting), accumulated per CTA in shared memory
            const int slot0w = __shfl_sync(FULL, slot_first, 0);
            if (__all_sync(FULL, same && (slot_first == slot0w || slot_first < 0))) {
                const u32 ds = warp_sum(dsum), vc = warp_sum(vcnt);
                if (lane == 0 && slot0w >= 0) {
                    atomicAdd(&sm.acc_edges[slot0w], ds);
                    atomicAdd(&sm.acc_verts[slot0w], vc);
                }
            } else {
#pragma unroll
                for (int q = 0; q < P2_UA; ++q)
                    if (e[q] != ~0ull) {
                        u32 d = entry_deg24(e[q]);
                        if (d == DEG_SAT) d = (u32)__ldg(&p.deg[entry_vertex(e[q])]);
                        atomicAdd(&sm.acc_edges[rs[q]], d);
                        atomicAdd(&sm.acc_verts[rs[q]], 1u);
                    }
            }
        }
        // shared counters can exceed 32 bits over a long launch: move them to the global 64-bit ones when they grow
        __syncthreads();
        if (threadIdx.x < (u32)k && (sm.acc_edges[threadIdx.x] > 0x40000000u || sm.acc_verts[threadIdx.x] > 0x40000000u)) {
            atomicAdd(&p.edges[a.slot0 + threadIdx.x], (u64)sm.acc_edges[threadIdx.x]);
            atomicAdd(&p.vertices[a.slot0 + threadIdx.x], (u64)sm.acc_verts[threadIdx.x]);
            sm.acc_edges[threadIdx.x] = 0;
            sm.acc_verts[threadIdx.x] = 0;
        }
        grid.sync();
        if (c == 0 && threadIdx.x < (u32)k) a.force[a.slot0 + threadIdx.x] = 0; // every CTA read it before the barrier above
        if (c == 0 && threadIdx.x == 0 && p.trace && level < p.trace_cap) {
            u64 t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            p.trace[4 * level + 3] = t;
        }
        // ---- phase B.  Hubs first: the listed entries are staged CTA-wide (one per thread), their edges form one line and every
        // CTA takes one slice of its 32*P2_UB-edge steps.  The steps and the CTA's 32-entry groups (handed out block-cyclically:
        // the degree mix of neighbouring frontier entries is correlated) then form ONE work list that the warps drain on their
        // own -- no CTA barrier below this point, a warp prefetches the entries of its next group while it scatters the current one.
        const u32 nh = min(*((volatile u32*)&a.hub_cnt[hp]), P2_HUB_CAP);
        u32 hub_steps = 0, hub_s0 = 0, EH = 0;
        if (nh) {
            u32 d = 0;
            if (threadIdx.x < nh) {
                const u32 i = __ldcg(&a.hub[hp * P2_HUB_CAP + threadIdx.x]);
                const int rsl = p2_slot_of(sm, k, i);
                const u64 e = __ldcg(p2_front(p, a.slot0 + rsl, sm.par[rsl]) + (i - sm.fbase[rsl]));
                const double r = -__ldcg(&a.rv[i]);
                const u64 beg = (u64)__ldcg(&((const OffT*)a.begs)[i]);
                d = entry_deg24(e);
                if (d == DEG_SAT) d = (u32)__ldg(&p.deg[entry_vertex(e)]);
                sm.hb_beg32[threadIdx.x] = (u32)beg;
                sm.hb_beghi[threadIdx.x] = (u32)(beg >> 32);
                sm.hb_inc[threadIdx.x] = ((1.0 - p.alpha) * r) / (double)d;
                sm.hb_slot[threadIdx.x] = rsl;
            }
            const u32 incl = warp_incl_scan(d);
            if (lane == 31) sm.hb_wtot[w] = incl;
            __syncthreads();
            const u32 wt = sm.hb_wtot[lane];
            const u32 wti = warp_incl_scan(wt);
            const u32 before = __shfl_sync(FULL, wti, w > 0 ? w - 1 : 0);
            EH = __shfl_sync(FULL, wti, 31);
            sm.hb_off[threadIdx.x] = (w > 0 ? before : 0u) + incl - d;
            if (threadIdx.x == P2_THREADS - 1) sm.hb_off[P2_HUB_SM] = EH;
            __syncthreads();
            const u32 nsteps = (EH + WARP * P2_UB - 1) / (WARP * P2_UB);
            hub_s0 = (u32)(((u64)nsteps * c) / G);
            hub_steps = (u32)(((u64)nsteps * (c + 1)) / G) - hub_s0;
        }
        // this CTA's groups: global group j = ((i / P2_CYC) * G + c) * P2_CYC + i % P2_CYC for local index i
        const u32 NG = (nf + WARP - 1) / WARP;
        const u32 nblk = (NG + P2_CYC - 1) / P2_CYC;                      // blocks of P2_CYC groups
        const u32 myblk = nblk > c ? (nblk - c + G - 1) / G : 0u;          // blocks c, c + G, ...
        const u32 items = hub_steps + myblk * P2_CYC;                      // (trailing groups past NG are skipped below)
        auto fetch = [&]() -> u32 {
            u32 it = 0;
            if (lane == 0) it = atomicAdd(&sm.next_group, 1u);
            return __shfl_sync(FULL, it, 0);
        };
        // per-lane registers of a prefetched group
        u64 pf_e = 0;
        OffT pf_beg = 0;
        double pf_r = 0.0;
        u32 pf_i = 0;
        bool pf_valid = false;
        auto prefetch_group = [&](u32 it) { // it: item index >= hub_steps
            const u32 li = it - hub_steps;
            const u32 gj = ((li / P2_CYC) * G + c) * P2_CYC + li % P2_CYC;
            pf_i = gj * WARP + lane;
            pf_valid = it < items && gj < NG && pf_i < nf;
            if (pf_valid) {
                const int rsl = p2_slot_of(sm, k, pf_i);
                pf_e = __ldcg(p2_front(p, a.slot0 + rsl, sm.par[rsl]) + (pf_i - sm.fbase[rsl]));
                pf_r = __ldcg(&a.rv[pf_i]);
                pf_beg = __ldcg(&((const OffT*)a.begs)[pf_i]);
            }
        };
        u32 item = fetch();
        if (item >= hub_steps && item < items) prefetch_group(item);
        while (item < items) {
            const u32 nxt = fetch();
            if (item < hub_steps) {
                if (nxt >= hub_steps && nxt < items) prefetch_group(nxt);
                const u32 x0 = (hub_s0 + item) * (WARP * P2_UB);
                p2_scatter<OffT, true>(a, g, sm, ctl, w, x0, min(EH, x0 + WARP * P2_UB), wq, wq_slot, pol_keep, pol_stream, nh);
            } else {
                // consume the prefetched registers, then start the next group's loads before scattering this one
                u32 cnt = 0;
                int rsl = 0;
                double inc = 0.0;
                OffT beg = 0;
                bool dang = false;
                if (pf_valid) {
                    rsl = p2_slot_of(sm, k, pf_i);
                    u32 d = entry_deg24(pf_e);
                    if (d == DEG_SAT) d = (u32)__ldg(&p.deg[entry_vertex(pf_e)]);
                    if (pf_r > 0.0) { // < 0: listed as a hub, handled through the hub steps
                        dang = d == 0;
                        beg = pf_beg;
                        inc = dang ? pf_r * (1.0 - p.alpha) : ((1.0 - p.alpha) * pf_r) / (double)d;
                        cnt = dang ? 1u : d;
                    }
                }
                if (nxt >= hub_steps && nxt < items) prefetch_group(nxt);
                else pf_valid = false;
                const u32 incl = warp_incl_scan(cnt);
                const u32 T = __shfl_sync(FULL, incl, 31);
                sm.w_off[w][lane] = incl - cnt;
                if (lane == 31) sm.w_off[w][32] = incl;
                sm.w_beg32[w][lane] = (u32)(u64)beg;
                sm.w_beghi[w][lane] = (u32)((u64)beg >> 32);
                sm.w_inc[w][lane] = inc;
                sm.w_slot[w][lane] = dang ? ~rsl : rsl;
                __syncwarp();
                if (T) p2_scatter<OffT, false>(a, g, sm, ctl, w, 0u, T, wq, wq_slot, pol_keep, pol_stream);
                __syncwarp();
            }
            item = nxt;
        }
        if (wq) {
            p2_flush(a, sm, ctl, sm.wqueue[w], wq, wq_slot);
            wq = 0;
        }
        if (p.trace && level < p.trace_cap && c == 0) { // when CTA 0's last warp is done with phase B (before the barrier)
            u64 t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (lane == 0) atomicMax(&p.trace[4 * level + 2], t);
        }
        grid.sync();
        if (threadIdx.x < P2_MAX_SUB) {
            sm.logpos[threadIdx.x] += sm.cnt[threadIdx.x];
            sm.par[threadIdx.x] ^= 1u;
        }
        __syncthreads();
    }
    // ---- hand the frontier state to the next kernel
    if (threadIdx.x < (u32)k) {
        const int t = threadIdx.x, as = a.slot0 + t;
        if (sm.acc_edges[t]) atomicAdd(&p.edges[as], (u64)sm.acc_edges[t]);
        if (sm.acc_verts[t]) atomicAdd(&p.vertices[as], (u64)sm.acc_verts[t]);
        if (c == 0) {
            ctl->par[as] = sm.par[t];
            ctl->fcount[sm.par[t] ^ 1u][as] = 0;
            if (p.log_v) p.log_cur[as] = sm.logpos[t];
        }
    }
    if (c == 0 && threadIdx.x == 0) {
        ctl->levels_run = levels_run;
        a.hub_cnt[0] = a.hub_cnt[1] = 0; // invariant between launches
    }
}

// -------------------------------------------------------------------------------------------------------------------
// Tail kernel: one CTA per slot runs every level whose frontier has <= tail_nf entries and <= tail_e edges with CTA-local
// barriers only.  The frontier lives in shared memory between levels (and is mirrored to the slot's global segment, so the
// kernel can hand over at any level).  Same rule, same schedule: phase A of a level completes (barrier) before phase B starts.
// -------------------------------------------------------------------------------------------------------------------
struct TailSmem {
    u64 front[2][TAIL_NF_CAP];
    u32 off[TAIL_NF_CAP + 1];
    u32 beglo[TAIL_NF_CAP], beghi[TAIL_NF_CAP];
    double inc[TAIL_NF_CAP];
    u32 wtot[TAIL_THREADS / WARP];
    u32 nnext;
    u32 total;
    u32 ndang;
};

template <typename OffT>
__global__ void __launch_bounds__(TAIL_THREADS, 1) push_tail_kernel(Push2Args a, CsrView<OffT> g) {
    extern __shared__ __align__(16) unsigned char tail_smem_raw[];
    TailSmem& sm = *reinterpret_cast<TailSmem*>(tail_smem_raw);
    const PushArgs& p = a.p;
    PushCtl* ctl = p.ctl;
    const int as = a.slot0 + (int)blockIdx.x;
    const int lane = lane_id(), w = threadIdx.x >> 5;
    u32 par = ctl->par[as];
    u32 nf = ctl->fcount[par][as];
    if (nf == 0) {
        if (threadIdx.x == 0) a.left[as] = 0;
        return;
    }
    const double rmax = p.rmax[as];
    const int32_t source = p.source[as];
    const u32 tail_nf = min(a.tail_nf, TAIL_NF_CAP);
    double* residue = p.residue + (size_t)as * p.n;
    u64* fb[2] = {p.front0 + (size_t)as * p.n, p.front1 + (size_t)as * p.n};
    const u32 dmax = p.colx ? (0xffffffffu >> p.deg_shift) : 0u, idmask = p.colx ? ((1u << p.deg_shift) - 1u) : 0xffffffffu;
    const int32_t* __restrict__ colp = p.colx ? p.colx : g.col;
    u32 logpos = p.log_v ? p.log_cur[as] : 0u;
    u64 acc_edges = 0, acc_verts = 0, acc_levels = 0;
    bool in_smem = false; // the current frontier is in sm.front[par]
    u32 level = 0;
    for (;; ++level) {
        if (nf == 0 || nf > tail_nf) break;
        if (level >= p.max_levels) {
            if (threadIdx.x == 0) *a.err = 1;
            break;
        }
        // entries, degrees, edge offsets (nothing is modified before the level is known to fit)
        if (threadIdx.x == 0) sm.ndang = 0;
        __syncthreads();
        u32 mydang = 0;
        for (u32 i = threadIdx.x; i < nf; i += TAIL_THREADS) {
            const u64 e = in_smem ? sm.front[par][i] : __ldcg(&fb[par][i]);
            if (!in_smem) sm.front[par][i] = e;
            u32 d = entry_deg24(e);
            if (d == DEG_SAT) d = (u32)__ldg(&p.deg[entry_vertex(e)]);
            sm.off[i] = d ? d : 1u; // a dangling vertex owns one pseudo-edge back to the source
            mydang += d == 0;
        }
        mydang = warp_sum(mydang);
        if (lane == 0 && mydang) atomicAdd(&sm.ndang, mydang);
        __syncthreads();
        // exclusive scan of sm.off[0..nf): thread t owns the contiguous run [t*per, (t+1)*per)
        const u32 per = (nf + TAIL_THREADS - 1) / TAIL_THREADS;
        const u32 r0 = min(nf, threadIdx.x * per), r1 = min(nf, r0 + per);
        u32 run = 0;
        for (u32 i = r0; i < r1; ++i) run += sm.off[i];
        const u32 incl = warp_incl_scan(run);
        if (lane == 31) sm.wtot[w] = incl;
        __syncthreads();
        if (w == 0) {
            const u32 t = sm.wtot[lane];
            const u32 ti = warp_incl_scan(t);
            sm.wtot[lane] = ti - t;
            if (lane == 31) sm.total = ti;
        }
        __syncthreads();
        const u32 E = sm.total;
        if (E > a.tail_e) { // too many edges for one CTA: the sub-wave kernel runs this level
            if (threadIdx.x == 0) a.force[as] = 1;
            break;
        }
        u32 base = sm.wtot[w] + incl - run;
        for (u32 i = r0; i < r1; ++i) {
            const u32 d = sm.off[i];
            sm.off[i] = base;
            base += d;
        }
        if (threadIdx.x == 0) {
            sm.off[nf] = E;
            sm.nnext = 0;
        }
        // phase A
        for (u32 i = threadIdx.x; i < nf; i += TAIL_THREADS) {
            const u64 e = sm.front[par][i];
            const int32_t v = entry_vertex(e);
            const double r = __longlong_as_double((long long)atomicExch((unsigned long long*)&residue[v], 0ull));
            u32 d = entry_deg24(e);
            if (d == DEG_SAT) d = (u32)__ldg(&p.deg[v]);
            const OffT beg = d ? g.ptr[v] : (OffT)0;
            const u32 lp = logpos + i;
            if (p.log_v && lp < p.log_cap) {
                const size_t li = (size_t)as * p.log_cap + lp;
                __stcs(&p.log_v[li], v);
                __stcs(&p.log_r[li], r);
            } else {
                const size_t gi = (size_t)as * p.n + v;
                p.reserve[gi] = __ldcg(&p.reserve[gi]) + r * p.alpha;
            }
            sm.inc[i] = d ? ((1.0 - p.alpha) * r) / (double)d : r * (1.0 - p.alpha);
            sm.beglo[i] = (u32)(u64)beg;
            sm.beghi[i] = d ? (u32)((u64)beg >> 32) : 0xffffffffu; // 0xffffffff: dangling (one pseudo-edge to the source)
        }
        __syncthreads();
        // phase B
        const u32 nb = par ^ 1u;
        for (u32 x0 = 0; x0 < E; x0 += TAIL_THREADS) {
            const u32 x = x0 + threadIdx.x;
            bool cross = false;
            u64 entry = 0;
            if (x < E) {
                u32 lo = 0, hi = nf; // largest o with off[o] <= x
                while (hi - lo > 1) {
                    const u32 mid = (lo + hi) >> 1;
                    if (sm.off[mid] <= x) lo = mid;
                    else hi = mid;
                }
                const double inc = sm.inc[lo];
                int32_t u;
                u32 dcode = dmax;
                if (sm.beghi[lo] == 0xffffffffu) {
                    u = source;
                } else {
                    const OffT beg = (OffT)(((u64)sm.beghi[lo] << 32) | sm.beglo[lo]);
                    const u32 raw = (u32)__ldcs(&colp[beg + (OffT)(x - sm.off[lo])]);
                    u = (int32_t)(raw & idmask);
                    if (p.colx) dcode = raw >> p.deg_shift;
                }
                const double old = atomicAdd(&residue[u], inc);
                const int32_t du = dcode != dmax ? (int32_t)dcode : __ldg(&p.deg[u]);
                const double thr = rmax * (double)du;
                cross = du ? (old < thr && old + inc >= thr) : (old == 0.0);
                entry = make_entry(as, (u32)du, u);
            }
            const u32 m = __ballot_sync(FULL, cross);
            if (m) {
                u32 b = 0;
                if (lane == 0) b = atomicAdd(&sm.nnext, (u32)__popc(m));
                b = __shfl_sync(FULL, b, 0);
                if (cross) {
                    const u32 pos = b + __popc(m & lanemask_lt());
                    fb[nb][pos] = entry;                       // the global segment always holds the frontier
                    if (pos < TAIL_NF_CAP) sm.front[nb][pos] = entry;
                }
            }
        }
        __syncthreads();
        acc_edges += E - sm.ndang; // the statistics count real out-edges only
        acc_verts += nf;
        acc_levels += 1;
        logpos += nf;
        const u32 nn = sm.nnext;
        __syncthreads();
        par = nb;
        nf = nn;
        in_smem = nn <= TAIL_NF_CAP;
    }
    if (threadIdx.x == 0) {
        ctl->par[as] = par;
        ctl->fcount[par][as] = nf;
        ctl->fcount[par ^ 1u][as] = 0;
        a.left[as] = nf;
        if (p.log_v) p.log_cur[as] = logpos;
        if (acc_verts) {
            p.edges[as] += acc_edges;
            p.vertices[as] += acc_verts;
            p.levels[as] += acc_levels;
        }
    }
}

} // namespace fora
