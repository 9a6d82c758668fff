// fora_b200/csrc/push3.cuh -- forward push, third generation: the grid sweeps the query slots in LOCKSTEP (sm_100a).
//
// Same per-vertex rule and the same level-synchronous schedule as push.cuh (/root/reference/algo.h:954-1093; the parity tests
// compare against a CPU restatement of exactly this schedule).  What changes is the order in which the memory system sees the
// scatters of a level, and how the next frontier is found.
//
// Measured (scripts/ubench_sweep*.cu, profiles/r2g_ubench_sweep*.txt): fp64 adds into S = 48 dense 39 MB vectors, 4 M per vector,
//   * slot-major but unsynchronised (first-generation kernel: CTAs drift over several slots inside a level)   ATOM  31 G/s
//   * a grid barrier after every vector (at most one or two vectors live in the 126 MB L2)                     ATOM  83 G/s
//   * the same with RED (no return value)                                                                       RED  120 G/s
//   * RED plus a dense scan of the vector (the next frontier found without return values)                           88 G/s
// A dirty sector that is evicted and re-fetched between two touches costs far more than its 64 bytes (the DRAM
// read-modify-write rate caps at ~20 G sectors/s), and without a barrier a wave of 48 slots keeps 3+ vectors live.  So:
//
//   * A level is cut into GROUPS of consecutive slots.  A slot whose frontier is large is a group of its own and runs in DENSE
//     mode; the other slots are grouped while their estimated scatter footprint fits one vector's worth of L2 (small levels:
//     all slots in one group, exactly the first-generation schedule).
//   * The grid works through the groups in order with one grid barrier per group.  Barrier interval j does phase A of group j
//     (exchange the frontier's residues, compute the increments) and phase B of group j - 1 (the scatters); a dense group is
//     scanned right behind its phase B, after one more barrier.
//   * Phase B is warp-autonomous: a warp claims 64 frontier entries (16 claim queues per interval, CTAs start on different
//     queues and move on when theirs is empty: dynamic balance without a hot cursor), expands 32 of them at a time by an
//     exclusive scan of their out-degrees in registers, finds every edge's owner by a 5-step shuffle search and keeps P3_UB
//     column loads and adds in flight per lane.  No edge-offset scan, no tile search, no CTA barrier.
//       - sparse mode: fp64 atomics WITH return; a vertex joins the next level exactly when the returned old value shows the
//         add crossed rmax*d_out (as in push.cuh); crossing vertices go through per-warp queues.
//       - dense mode: fp64 RED, nobody waits for a return value.  After the barrier the slot's residue vector (still in the L2)
//         is scanned once: every vertex at or above its threshold IS the next frontier (phase A zeroed the previous one, so
//         nothing else can be above), and the scan does that frontier's phase A on the spot (zero, credit log, increment,
//         adjacency start).  The next frontier comes out sorted by vertex id, so its column reads walk the CSR forward.
//   * Vertices with more than hub_deg out-edges are cut into pieces when their entry is made; the pieces are claimed warp by
//     warp after the queues, so a hub is spread over the whole grid.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "push.cuh"

namespace fora {

constexpr int P3_THREADS = 512;
constexpr int P3_WARPS = P3_THREADS / WARP;
#ifndef CFG_P3_UB
#define CFG_P3_UB 4
#endif
constexpr int P3_UB = CFG_P3_UB;          // edges in flight per lane in phase B, sparse mode (atomics with return)
#ifndef CFG_P3_UBD
#define CFG_P3_UBD 8
#endif
constexpr int P3_UBD = CFG_P3_UBD;        // edges in flight per lane in phase B, dense mode (column loads, then REDs)
#ifndef CFG_P3_UA
#define CFG_P3_UA 4
#endif
constexpr int P3_UA = CFG_P3_UA;          // frontier entries in flight per thread in phase A
#ifndef CFG_P3_PIECE
#define CFG_P3_PIECE 32
#endif
constexpr u32 P3_PIECE = CFG_P3_PIECE;    // frontier entries per warp claim (one batch: a claim is at most 32 * hub_deg edges)
constexpr int P3_NQ = 16;                 // claim queues per interval
constexpr u32 P3_HUB_DEG = 64;            // a frontier vertex with more out-edges is cut into pieces
constexpr u32 P3_HUB_PIECE = 256;         // edges per hub piece
constexpr int P3_NQH = 8;                 // claim cursors over a slot's hub pieces
constexpr u32 P3_HUB_CHUNK = 4;           // hub pieces per claim
constexpr int P3_WQ = 256;                // per-warp queue of crossing vertices
#ifndef CFG_P3_AWARPS
#define CFG_P3_AWARPS 4
#endif
constexpr int P3_AWARPS = CFG_P3_AWARPS;  // warps per CTA that run phase A of the next group while the others scatter
constexpr int P3_SCAN_K = 32;             // dense scan: vertices per thread and tile (tile = 32 * 512 vertices)

struct P3Ctl {
    u32 cursor[3][P3_NQ];        // claim cursors over a group's batches of entries, rotating per barrier interval
    u32 hcursor[3][P3_NQH];      // ... over its hub pieces
    u32 hubcnt[3][MAX_SLOTS];    // hub pieces listed for the slot's frontier of level L: set L % 3 (rotates like PushCtl::fcount)
};

struct P3Args {
    P3Ctl* c;
    u64* hub_list;       // [2][slots][hub_cap] by level parity: (index inside the slot's frontier << 24) | piece number
    u32 hub_cap;         // per slot
    u32 est_deg;         // edges assumed per frontier vertex when a level is cut into groups
    u32 budget_sectors;  // a sparse group's estimated footprint in 32-byte sectors (default: one residue vector = n / 4)
    u32 hub_deg;         // a frontier vertex with more out-edges is cut into pieces of hub_piece edges (P3_HUB_DEG / P3_HUB_PIECE;
    u32 hub_piece;       // smaller values are a test hook)
    u32 dense_min;       // frontier entries from which a slot-level runs in dense mode (0xffffffff: never)
    u32 debug_mode;      // ablation (wrong answers, timing only), dense mode: bit 0 no REDs, bit 1 no column loads, bit 2 no expansion at all
    u32 debug_skip_hot;  // ablation (wrong answers, timing only): dense-mode adds to vertices below this id are dropped
    u32* beg32;          // [slots*n] per frontier entry: start of the adjacency list (32-bit offsets only; otherwise g.ptr is read)
};

struct P3Smem {
    u64 wqueue[P3_WARPS][P3_WQ]; // crossing vertices, one private queue per warp
    double rmax[MAX_SLOTS];
    int32_t source[MAX_SLOTS];
    u32 cnt_edges[MAX_SLOTS], cnt_verts[MAX_SLOTS];
    u32 fbase[MAX_SLOTS + 1];    // the level's frontier = the slots' segments concatenated in slot order
    u32 logbase[MAX_SLOTS];
    u32 prevcnt[MAX_SLOTS];
    int gstart[MAX_SLOTS + 1];   // group g = slots [gstart[g], gstart[g + 1])
    unsigned char dense[MAX_SLOTS]; // this level: the slot is a dense group
    unsigned char adone[MAX_SLOTS]; // this level: the slot's frontier was made by a scan (phase A already done)
    unsigned char inlevel[MAX_SLOTS]; // this level: the slot has a frontier
    int ng;
    u32 qdone;                   // bit q: claim queue q of the current interval is empty
    // dense scan
    u32 sc_cnt[P3_SCAN_K * P3_WARPS];
    u32 sc_wtot[P3_WARPS];
    u32 sc_base;
    u32 sc_list[P3_SCAN_K * P3_THREADS]; // hit vertices of the tile, in vertex order
};

// sum over vertices with more than hub_deg out-edges of ceil(d / piece): capacity of a slot's hub piece list (push3.cuh)
__global__ void hub_pieces_kernel(int32_t n, const int32_t* __restrict__ deg, u32 hub_deg, u32 piece, u64* out) {
    u64 s = 0;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
        const u32 d = (u32)deg[v];
        if (d > hub_deg) s += (d + piece - 1) / piece;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// slot of frontier index i among slots [s0, s1) (fbase[s0] <= i < fbase[s1])
__device__ __forceinline__ int p3_slot_of(const P3Smem& sm, int s0, int s1, u32 i) {
    int lo = s0, hi = s1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (sm.fbase[mid] <= i) lo = mid;
        else hi = mid;
    }
    return lo;
}

// streamed once per level: keep them out of the way of the residue vector the grid is working on
__device__ __forceinline__ u64 ld_u64_stream(const u64* addr, u64 policy) {
    u64 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(addr), "l"(policy));
    return v;
}
__device__ __forceinline__ double ld_f64_stream(const double* addr, u64 policy) {
    double v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(addr), "l"(policy));
    return v;
}
__device__ __forceinline__ u32 ld_u32_stream(const u32* addr, u64 policy) {
    u32 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(addr), "l"(policy));
    return v;
}
__device__ __forceinline__ void st_u64_stream(u64* addr, u64 v, u64 policy) {
    asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(addr), "l"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ void st_f64_stream(double* addr, double v, u64 policy) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(addr), "d"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ void st_u32_stream(u32* addr, u32 v, u64 policy) {
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(addr), "r"(v), "l"(policy) : "memory");
}

// list the pieces of a hub entry (index j inside slot s's frontier of level `lvl`)
__device__ __forceinline__ void p3_list_hub(const PushArgs& a, const P3Args& x, int s, u32 lvl, u32 j, u32 d) {
    const u32 np = (d + x.hub_piece - 1) / x.hub_piece;
    const u32 hb = atomicAdd(&x.c->hubcnt[lvl % 3][s], np);
    u64* list = x.hub_list + ((size_t)(lvl & 1) * a.slots + s) * x.hub_cap;
    for (u32 p = 0; p < np; ++p) {
        if (hb + p < x.hub_cap) list[hb + p] = ((u64)j << 24) | p;
        else if (a.err) *a.err = 2;
    }
}

// ---- phase A of slots [s0, s1): CTA `rank` takes a contiguous share of their frontier entries ----------------------------------
template <typename OffT>
__device__ __forceinline__ void p3_phase_a(const PushArgs& a, const CsrView<OffT>& g, const P3Args& x, P3Smem& sm, const u64* cur,
                                           int s0, int s1, u32 rank, u32 count, u32 level, u32 tid, u32 nthreads) {
    const int lane = lane_id();
    const u32 gb = sm.fbase[s0], ge = sm.fbase[s1];
    const u32 len = ge - gb;
    const u32 lo = gb + (u32)(((u64)len * rank) / count), hi = gb + (u32)(((u64)len * (rank + 1)) / count);
    if (lo >= hi) return;
    for (int s = p3_slot_of(sm, s0, s1, lo); s < s1 && sm.fbase[s] < hi; ++s) {
        const u32 b = max(lo, sm.fbase[s]), e_ = min(hi, sm.fbase[s + 1]);
        if (b >= e_ || sm.adone[s]) continue;
        const u64* seg = cur + (size_t)s * a.n;
        double* res = a.residue + (size_t)s * a.n;
        double* incs = a.inc + (size_t)s * a.n;
        u32* begs = x.beg32 + (size_t)s * a.n;
        const u32 fb = sm.fbase[s], lb = sm.logbase[s];
        u32 dsum_w = 0, vcnt_w = 0;
        for (u32 i0 = b; i0 < e_; i0 += nthreads * P3_UA) {
            u64 e[P3_UA];
            double r[P3_UA];
            u32 d[P3_UA];
            OffT beg[P3_UA];
#pragma unroll
            for (int k = 0; k < P3_UA; ++k) {
                const u32 i = i0 + k * nthreads + tid;
                e[k] = i < e_ ? __ldcg(&seg[i - fb]) : ~0ull;
            }
#pragma unroll
            for (int k = 0; k < P3_UA; ++k) {
                r[k] = 0.0; d[k] = 0; beg[k] = 0;
                if (e[k] != ~0ull) {
                    const u32 v = (u32)e[k];
                    // read + zero in ONE L2 operation (a load followed by a plain store of 0 takes a slow path when the store
                    // reaches the L2 while the sector's fill is pending, push.cuh)
                    r[k] = __longlong_as_double((long long)atomicExch((unsigned long long*)&res[v], 0ull));
                    d[k] = entry_deg24(e[k]);
                    if (d[k] == DEG_SAT) d[k] = (u32)__ldg(&a.deg[v]);
                    if (sizeof(OffT) == 4) beg[k] = g.ptr[v];
                }
            }
            u32 dsum = 0, vcnt = 0;
#pragma unroll
            for (int k = 0; k < P3_UA; ++k) {
                if (e[k] == ~0ull) continue;
                const u32 j = i0 + k * nthreads + tid - fb; // index inside the slot's frontier
                const u32 v = (u32)e[k];
                const u32 lp = lb + j;
                if (a.log_v && lp < a.log_cap) {
                    const size_t li = (size_t)s * a.log_cap + lp;
                    a.log_v[li] = (int32_t)v;
                    a.log_r[li] = r[k];
                } else { // no room (or no log): direct update of the reserve
                    double* rp = &a.reserve[(size_t)s * a.n + v];
                    *rp = __ldcg(rp) + r[k] * a.alpha;
                }
                incs[j] = d[k] ? ((1.0 - a.alpha) * r[k]) / (double)d[k] : r[k] * (1.0 - a.alpha);
                if (sizeof(OffT) == 4) begs[j] = (u32)beg[k];
                if (d[k] > x.hub_deg) p3_list_hub(a, x, s, level, j, d[k]); // phase B's per-entry path skips it
                dsum += d[k];
                ++vcnt;
            }
            dsum_w += dsum;
            vcnt_w += vcnt;
        }
        // per-slot work counters (cost model of --balanced, roofline accounting): shared memory first
        const u32 ds = warp_sum(dsum_w), vc = warp_sum(vcnt_w);
        if (lane == 0 && vc) {
            atomicAdd(&sm.cnt_edges[s], ds);
            atomicAdd(&sm.cnt_verts[s], vc);
        }
    }
}

// per-warp state of the crossing queue
struct P3Queue {
    u32 wq;       // entries queued (warp-uniform)
    int wq_slot;  // the slot they belong to
};

// One scatter step of a warp: every lane holds up to UB edges (column index, increment, slot; dang: the pseudo-edge of a
// dangling vertex back to the slot's source).  Issues all column loads, then all adds; sparse mode then tests the thresholds.
template <typename OffT, bool DENSE, int UB>
__device__ __forceinline__ void p3_scatter(const PushArgs& a, const CsrView<OffT>& g, const P3Args& x, P3Smem& sm, const bool (&ok)[UB],
                                           const OffT (&cidx)[UB], const double (&inc)[UB], const int (&slot)[UB],
                                           const bool (&dang)[UB], u64* myq, P3Queue& q, u64* nxt, u32* nxt_count, u64 pol_keep,
                                           u64 pol_stream) {
    const u32 dmax = a.colx ? (0xffffffffu >> a.deg_shift) : 0u, idmask = a.colx ? ((1u << a.deg_shift) - 1u) : 0xffffffffu;
    const int32_t* __restrict__ colp = a.colx ? a.colx : g.col;
    if (!DENSE && q.wq > P3_WQ - WARP * UB) { // make room: one global atomic per flush
        push_flush_warp(a, myq, q.wq, q.wq_slot, nxt, nxt_count);
        q.wq = 0;
    }
    int32_t u[UB];
    u32 dcode[UB];
#pragma unroll
    for (int k = 0; k < UB; ++k) {
        u[k] = 0; dcode[k] = dmax;
        if (ok[k]) {
            if (dang[k]) {
                u[k] = sm.source[slot[k]];
            } else if (DENSE && (x.debug_mode & 2u)) {
                u[k] = (int32_t)(((u64)((u32)cidx[k] * 2654435761u) * (u32)a.n) >> 32);
            } else {
                const int32_t* cp = &colp[cidx[k]];
                const u32 raw = (u32)(a.l2_hints ? ld_col_stream(cp, pol_stream) : __ldcs(cp));
                u[k] = (int32_t)(raw & idmask);
                if (a.colx) dcode[k] = raw >> a.deg_shift;
            }
        }
    }
    if (DENSE) {
#pragma unroll
        for (int k = 0; k < UB; ++k)
            if (ok[k] && (u32)u[k] >= x.debug_skip_hot && !(x.debug_mode & 1u)) {
                double* rp = &a.residue[(size_t)slot[k] * a.n + u[k]];
                if (a.l2_hints) red_add_f64_hint(rp, inc[k], pol_keep);
                else atomicAdd(rp, inc[k]);
            }
        return;
    }
    double old[UB];
    int32_t du[UB];
#pragma unroll
    for (int k = 0; k < UB; ++k) {
        du[k] = 0; old[k] = 0.0;
        if (ok[k]) {
            double* rp = &a.residue[(size_t)slot[k] * a.n + u[k]];
            old[k] = a.l2_hints ? atomic_add_f64_hint(rp, inc[k], pol_keep) : atomicAdd(rp, inc[k]);
            if (dcode[k] != dmax) du[k] = (int32_t)dcode[k];
            else du[k] = __ldg(&a.deg[u[k]]);
        }
    }
#pragma unroll
    for (int k = 0; k < UB; ++k) {
        bool cross = false;
        if (ok[k]) {
            const double nw = old[k] + inc[k];
            const double thr = sm.rmax[slot[k]] * (double)du[k];
            cross = du[k] ? (old[k] < thr && nw >= thr) : (old[k] == 0.0);
        }
        // the queue holds one slot at a time; a step rarely spans two (a claim at a slot boundary of a multi-slot group)
        u32 pending = __ballot_sync(FULL, cross);
        while (pending) {
            const int s0 = __shfl_sync(FULL, slot[k], __ffs(pending) - 1);
            if (q.wq && s0 != q.wq_slot) {
                push_flush_warp(a, myq, q.wq, q.wq_slot, nxt, nxt_count);
                q.wq = 0;
            }
            q.wq_slot = s0;
            const u32 m = __ballot_sync(FULL, cross && slot[k] == s0);
            if (cross && slot[k] == s0) myq[q.wq + __popc(m & lanemask_lt())] = make_entry(s0, (u32)du[k], u[k]);
            q.wq += __popc(m);
            pending &= ~m;
        }
    }
}

// Expansion of one warp batch: lane l holds a run of `cnt` consecutive column slots starting at `beg` (a frontier entry's
// adjacency list, a piece of a hub's list, or the single pseudo-edge of a dangling vertex), its increment and its slot.  The
// runs are laid end to end (exclusive scan in registers); edge p of the batch belongs to the largest lane whose offset is <= p
// (5-step shuffle search; lanes that contribute nothing tie with their successor and lose).
template <typename OffT, bool DENSE, int UB>
__device__ __forceinline__ void p3_expand(const PushArgs& a, const CsrView<OffT>& g, const P3Args& x, P3Smem& sm, u32 cnt, OffT beg, double inc,
                                          int slot, bool dg, u64* myq, P3Queue& q, u64* nxt, u32* nxt_count, u64 pol_keep,
                                          u64 pol_stream) {
    const int lane = lane_id();
    const u32 incl = warp_incl_scan(cnt);
    const u32 excl = incl - cnt;
    const u32 T = __shfl_sync(FULL, incl, 31);
    const u32 inc_lo = (u32)__double_as_longlong(inc), inc_hi = (u32)(__double_as_longlong(inc) >> 32);
    const int flags = (slot << 1) | (dg ? 1 : 0);
    for (u32 base = 0; base < T; base += WARP * UB) {
        bool ok[UB], dang[UB];
        OffT cidx[UB];
        double einc[UB];
        int eslot[UB];
#pragma unroll
        for (int k = 0; k < UB; ++k) {
            const u32 p = base + k * WARP + lane;
            ok[k] = p < T;
            int own = 0;
#pragma unroll
            for (int st = 16; st > 0; st >>= 1) {
                const u32 v = __shfl_sync(FULL, excl, own + st);
                if (v <= p) own += st;
            }
            const u32 oex = __shfl_sync(FULL, excl, own);
            const OffT obeg = __shfl_sync(FULL, beg, own);
            const u32 lo32 = __shfl_sync(FULL, inc_lo, own), hi32 = __shfl_sync(FULL, inc_hi, own);
            const int fl = __shfl_sync(FULL, flags, own);
            cidx[k] = obeg + (OffT)(p - oex);
            einc[k] = __longlong_as_double((long long)(((u64)hi32 << 32) | lo32));
            eslot[k] = fl >> 1;
            dang[k] = fl & 1;
        }
        p3_scatter<OffT, DENSE, UB>(a, g, x, sm, ok, cidx, einc, eslot, dang, myq, q, nxt, nxt_count, pol_keep, pol_stream);
    }
}

// Striped work pool: P items over nq cursors; the k-th claim on cursor q is item k * nq + q, so every cursor sees the same mix of
// items and no single address takes all the atomics.  A warp starts on its CTA's cursor and moves on cyclically; when a claim
// fails it looks at ALL cursors at once (one round trip), so running dry costs two memory latencies, not one per cursor.
struct P3Claim {
    u32* cursor;
    u32 nq, P;
    u32 qi;     // current cursor
    u32 alive;  // cursors this warp still believes non-empty (warp-uniform)
};
// claims `chunk` consecutive k on the current live cursor; returns k (items (k + i) * nq + qi, i < chunk, those < P) or ~0 when dry
__device__ __forceinline__ u32 p3_claim(P3Claim& c, u32 chunk, u32* qdone) {
    const int lane = lane_id();
    for (;;) {
        if (qdone) c.alive &= ~__shfl_sync(FULL, *(volatile u32*)qdone, 0); // what sibling warps found out
        if (!c.alive) return 0xffffffffu;
        const u32 hi_q = c.alive & ~((1u << c.qi) - 1u); // next live cursor at or after qi, cyclically
        c.qi = (u32)__ffs(hi_q ? hi_q : c.alive) - 1u;
        u32 k = 0;
        if (lane == 0) k = atomicAdd(&c.cursor[c.qi], chunk);
        k = __shfl_sync(FULL, k, 0);
        if ((u64)k * c.nq + c.qi < c.P) return k;
        bool live = false;
        if ((u32)lane < c.nq) live = (u64)(*(volatile u32*)&c.cursor[lane]) * c.nq + (u32)lane < c.P;
        const u32 m = __ballot_sync(FULL, live);
        c.alive &= m & ~(1u << c.qi);
        if (qdone && lane == 0) atomicOr(qdone, ~m & ((1u << c.nq) - 1u));
    }
}

// ---- phase B of slots [s0, s1) (at most 32 slots) --------------------------------------------------------------------------------
template <typename OffT, bool DENSE>
__device__ __forceinline__ void p3_phase_b(const PushArgs& a, const CsrView<OffT>& g, const P3Args& x, P3Smem& sm, const u64* cur,
                                           int s0, int s1, u32 set_b, u32 level, u64* nxt, u32* nxt_count) {
    constexpr int UB = DENSE ? P3_UBD : P3_UB;
    const int lane = lane_id(), w = threadIdx.x >> 5;
    const u32 gb = sm.fbase[s0], ge = sm.fbase[s1];
    const u32 len = ge - gb;
    if (len == 0) return;
    const u64 pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
    u64* myq = sm.wqueue[w];
    P3Queue q{0u, 0};
    // the group's hub pieces form one pool (the slots' lists laid end to end): counts first, they are needed last
    const u32 hcnt = s0 + lane < s1 ? min(*(volatile u32*)&x.c->hubcnt[level % 3][s0 + lane], x.hub_cap) : 0u;
    // batches of 32 consecutive entries, claimed from the back: a scan-made frontier is sorted by vertex id, and the relabelling puts
    // the vertices with many out-edges last -- the heavy batches go first, the light ones fill the end of the interval
    const u32 P = (len + P3_PIECE - 1) / P3_PIECE;
    P3Claim cl{x.c->cursor[set_b], min((u32)P3_NQ, max(1u, P / 8u)), P, 0u, 0u};
    cl.qi = blockIdx.x % cl.nq;
    cl.alive = (1u << cl.nq) - 1u;
#ifdef CFG_P3_STATIC
    const u32 gw = blockIdx.x * P3_WARPS + w, nw = gridDim.x * P3_WARPS;
    for (u32 it = gw; it < P; it += nw) {
        const u32 batch = P - 1u - it;
#else
    for (;;) {
        const u32 k = p3_claim(cl, 1u, &sm.qdone);
        if (k == 0xffffffffu) break;
        const u32 batch = P - 1u - (k * cl.nq + cl.qi);
#endif
      for (u32 sub = 0; sub < P3_PIECE; sub += WARP) {
        const u32 i = gb + batch * P3_PIECE + sub + lane; // one frontier entry per lane
        if (gb + batch * P3_PIECE + sub >= ge) break;
        int slot = s0;
        u32 cnt = 0; // edges this lane's entry contributes here (0: none / hub, 1 for a dangling vertex)
        bool dg = false;
        double inc = 0.0;
        OffT beg = 0;
        if (i < ge) {
            slot = s1 - s0 > 1 ? p3_slot_of(sm, s0, s1, i) : s0;
            const size_t si = (size_t)slot * a.n + (i - sm.fbase[slot]);
            const u64 e = ld_u64_stream(&cur[si], pol_stream);
            u32 d = entry_deg24(e);
            if (d == DEG_SAT) d = (u32)__ldg(&a.deg[(u32)e]);
            inc = ld_f64_stream(&a.inc[si], pol_stream);
            if (sizeof(OffT) == 4) beg = (OffT)ld_u32_stream(&x.beg32[si], pol_stream);
            else beg = g.ptr[(u32)e];
            dg = d == 0;
            cnt = dg ? 1u : (d > x.hub_deg ? 0u : d);
        }
        if (DENSE && (x.debug_mode & 4u)) { if (cnt == 0xfffffffu) *nxt_count = 1; continue; }
        p3_expand<OffT, DENSE, UB>(a, g, x, sm, cnt, beg, inc, slot, dg, myq, q, nxt, nxt_count, pol_keep, pol_stream);
      }
    }
    // hub pieces: P3_HUB_CHUNK pieces per claim, one per lane, expanded like a batch of entries
    const u32 hincl = warp_incl_scan(hcnt);
    const u32 hexcl = hincl - hcnt;
    const u32 HP = __shfl_sync(FULL, hincl, 31);
    if (HP) {
        P3Claim hcl{x.c->hcursor[set_b], min((u32)P3_NQH, max(1u, HP / (2u * P3_HUB_CHUNK))), HP, 0u, 0u};
        hcl.qi = blockIdx.x % hcl.nq;
        hcl.alive = (1u << hcl.nq) - 1u;
        for (;;) {
            const u32 hk = p3_claim(hcl, P3_HUB_CHUNK, nullptr);
            if (hk == 0xffffffffu) break;
            const u32 h = (hk + (u32)lane) * hcl.nq + hcl.qi; // lane's piece (lanes >= P3_HUB_CHUNK: none)
            const bool mine = (u32)lane < P3_HUB_CHUNK && h < HP;
            // slot of piece h: the largest lane (= slot - s0) whose offset is <= h
            int own = 0;
#pragma unroll
            for (int st = 16; st > 0; st >>= 1) {
                const u32 v = __shfl_sync(FULL, hexcl, own + st);
                if (v <= h) own += st;
            }
            const u32 hoff = __shfl_sync(FULL, hexcl, own);
            int slot = s0;
            u32 cnt = 0;
            double inc = 0.0;
            OffT beg = 0;
            if (mine) {
                slot = s0 + own;
                const u64 pc = __ldcg(&x.hub_list[((size_t)(level & 1) * a.slots + slot) * x.hub_cap + (h - hoff)]);
                const u32 j = (u32)(pc >> 24), piece = (u32)(pc & 0xffffffu);
                const size_t si = (size_t)slot * a.n + j;
                const u64 e = __ldcg(&cur[si]);
                u32 d = entry_deg24(e);
                if (d == DEG_SAT) d = (u32)__ldg(&a.deg[(u32)e]);
                inc = __ldcg(&a.inc[si]);
                const OffT b0 = sizeof(OffT) == 4 ? (OffT)__ldcg(&x.beg32[si]) : g.ptr[(u32)e];
                const u32 eb = piece * x.hub_piece;
                beg = b0 + (OffT)eb;
                cnt = min(d - eb, x.hub_piece);
            }
            p3_expand<OffT, DENSE, UB>(a, g, x, sm, cnt, beg, inc, slot, false, myq, q, nxt, nxt_count, pol_keep, pol_stream);
        }
    }
    if (!DENSE && q.wq) push_flush_warp(a, myq, q.wq, q.wq_slot, nxt, nxt_count);
}

// ---- dense scan of slot s after its dense phase B of level `level`: the vertices at or above their threshold are the frontier
// of level + 1; their phase A happens here.  CTA `rank` owns a contiguous range of the vector, handled in tiles of
// P3_SCAN_K * P3_THREADS vertices.  Pass 1 reads residue and out-degree of the whole tile, eight independent loads of each in
// flight per thread, and notes the hits (bit k of `mask` = vertex tile + k*512 + tid); a block scan of the 512 (k, warp) counts and
// ONE global atomic place the tile's hits in vertex order; the hit vertices are compacted into shared memory; pass 2 walks that
// list with all lanes busy and independent loads: entry, increment, adjacency start, credit log, residue zeroed.  Whole CTA,
// contains CTA barriers.
template <typename OffT>
__device__ __forceinline__ void p3_scan(const PushArgs& a, const CsrView<OffT>& g, const P3Args& x, P3Smem& sm, int s, u32 level,
                                        u32 rank, u32 count, u64* nxt, u32* nxt_count) {
    const int lane = lane_id(), w = threadIdx.x >> 5;
    const u32 n = (u32)a.n;
    const u32 per = (n + count - 1) / count;
    const u32 lo = min(n, rank * per), hi = min(n, lo + per);
    double* res = a.residue + (size_t)s * n;
    const double rm = sm.rmax[s];
    const u32 lb = sm.logbase[s] + sm.prevcnt[s]; // log position of the first entry of the slot's next frontier
    const u64 pol_stream = l2_policy_evict_first();
    u32 dsum_t = 0, vcnt_t = 0;
    for (u32 tb = lo; tb < hi; tb += P3_SCAN_K * P3_THREADS) {
        u32 mask = 0;
        for (int k0 = 0; k0 < P3_SCAN_K; k0 += 8) {
            double r[8];
            int32_t d[8];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const u32 v = tb + (k0 + kk) * P3_THREADS + threadIdx.x;
                r[kk] = v < hi ? __ldcg(&res[v]) : 0.0; // L2, never a stale L1 line: the REDs of other SMs just landed there
                d[kk] = v < hi ? __ldg(&a.deg[v]) : 1;  // unconditionally: one round trip instead of two
            }
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const bool hit = r[kk] > 0.0 && (d[kk] ? (r[kk] >= rm * (double)d[kk]) : true);
                const u32 bal = __ballot_sync(FULL, hit);
                if (hit) mask |= 1u << (k0 + kk);
                if (lane == 0) sm.sc_cnt[(k0 + kk) * P3_WARPS + w] = __popc(bal);
            }
        }
        __syncthreads();
        // exclusive scan of the 512 (k, warp) counts in vertex order: thread t owns count t
        const u32 c = sm.sc_cnt[threadIdx.x];
        const u32 incl = warp_incl_scan(c);
        if (lane == 31) sm.sc_wtot[w] = incl;
        __syncthreads();
        const u32 wt = lane < P3_WARPS ? sm.sc_wtot[lane] : 0u;
        const u32 wi = warp_incl_scan(wt);
        const u32 total = __shfl_sync(FULL, wi, P3_WARPS - 1);
        const u32 before = w ? __shfl_sync(FULL, wi, w - 1) : 0u;
        if (threadIdx.x == 0) sm.sc_base = total ? atomicAdd(&nxt_count[s], total) : 0u;
        sm.sc_cnt[threadIdx.x] = before + incl - c; // every thread rewrites only the count it read itself
        __syncthreads();
        if (total) { // warp-uniform (block-uniform)
            // compact the hit vertices, in vertex order
            for (int k = 0; k < P3_SCAN_K; ++k) {
                const bool hit = (mask >> k) & 1u;
                const u32 bal = __ballot_sync(FULL, hit);
                if (hit) sm.sc_list[sm.sc_cnt[k * P3_WARPS + w] + __popc(bal & lanemask_lt())] = tb + k * P3_THREADS + threadIdx.x;
            }
            __syncthreads();
            const u32 base = sm.sc_base;
            u64* seg = nxt + (size_t)s * n;
            double* incs = a.inc + (size_t)s * n;
            u32* begs = x.beg32 + (size_t)s * n;
            for (u32 h0 = 0; h0 < total; h0 += 4 * P3_THREADS) {
                u32 v[4], d[4];
                double r[4];
                OffT pb[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const u32 h = h0 + q * P3_THREADS + threadIdx.x;
                    v[q] = h < total ? sm.sc_list[h] : 0xffffffffu;
                    r[q] = 0.0; d[q] = 0; pb[q] = 0;
                    if (v[q] != 0xffffffffu) {
                        r[q] = __ldcg(&res[v[q]]);
                        d[q] = (u32)__ldg(&a.deg[v[q]]);
                        if (sizeof(OffT) == 4) pb[q] = g.ptr[v[q]];
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (v[q] == 0xffffffffu) continue;
                    const u32 j = base + h0 + q * P3_THREADS + threadIdx.x;
                    res[v[q]] = 0.0;
                    st_u64_stream(&seg[j], make_entry(s, d[q], (int32_t)v[q]), pol_stream);
                    const u32 lp = lb + j;
                    if (a.log_v && lp < a.log_cap) {
                        const size_t li = (size_t)s * a.log_cap + lp;
                        st_u32_stream((u32*)&a.log_v[li], v[q], pol_stream);
                        st_f64_stream(&a.log_r[li], r[q], pol_stream);
                    } else {
                        double* rp = &a.reserve[(size_t)s * n + v[q]];
                        *rp = __ldcg(rp) + r[q] * a.alpha;
                    }
                    st_f64_stream(&incs[j], d[q] ? ((1.0 - a.alpha) * r[q]) / (double)d[q] : r[q] * (1.0 - a.alpha), pol_stream);
                    if (sizeof(OffT) == 4) st_u32_stream(&begs[j], (u32)pb[q], pol_stream);
                    if (d[q] > x.hub_deg) p3_list_hub(a, x, s, level + 1, j, d[q]);
                    dsum_t += d[q];
                    ++vcnt_t;
                }
            }
        }
        __syncthreads(); // sc_cnt / sc_base / sc_list are rewritten by the next tile
    }
    const u32 ds = warp_sum(dsum_t), vc = warp_sum(vcnt_t);
    if (lane == 0 && vc) {
        atomicAdd(&sm.cnt_edges[s], ds);
        atomicAdd(&sm.cnt_verts[s], vc);
    }
}

template <typename OffT>
__global__ void __launch_bounds__(P3_THREADS, 2) push3_kernel(PushArgs a, CsrView<OffT> g, P3Args x) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char push3_smem_raw[];
    P3Smem& sm = *reinterpret_cast<P3Smem*>(push3_smem_raw);
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
    if (blockIdx.x == 0) // rotating state starts clean
        for (u32 i = threadIdx.x; i < sizeof(P3Ctl) / sizeof(u32); i += blockDim.x) reinterpret_cast<u32*>(x.c)[i] = 0;
    grid.sync();

    u32 t = 0; // barrier interval counter, identical in every CTA
    for (u32 level = 0;; ++level) {
        // the level's frontier: per-slot segments concatenated in slot order; cut into groups of consecutive slots
        if (threadIdx.x < WARP) {
            const int l = threadIdx.x;
            const u32 c0 = l < a.slots ? *((volatile u32*)&ctl->fcount[level % 3][l]) : 0u;
            const u32 c1 = l + WARP < a.slots ? *((volatile u32*)&ctl->fcount[level % 3][l + WARP]) : 0u;
            const u32 i0 = warp_incl_scan(c0);
            const u32 i1 = warp_incl_scan(c1) + __shfl_sync(FULL, i0, 31);
            if (l < a.slots) { sm.fbase[l] = i0 - c0; sm.logbase[l] += sm.prevcnt[l]; sm.prevcnt[l] = c0; }
            if (l + WARP < a.slots) { sm.fbase[l + WARP] = i1 - c1; sm.logbase[l + WARP] += sm.prevcnt[l + WARP]; sm.prevcnt[l + WARP] = c1; }
            if (l == 31) sm.fbase[a.slots] = i1;
            __syncwarp();
            if (l == 0) { // <= 64 slots: a serial pass is cheaper than being clever
                const u32 cap = max(1u, (u32)a.n / 4u);
                int ng = 0;
                u64 acc = 0;
                bool open = false; // a sparse group is being filled
                for (int s = 0; s < a.slots; ++s) {
                    const u32 c = sm.fbase[s + 1] - sm.fbase[s];
                    sm.adone[s] = sm.dense[s]; // a dense slot-level left its next frontier ready-made
                    sm.inlevel[s] = c > 0;
                    const bool dn = c >= x.dense_min && c > 0;
                    sm.dense[s] = dn;
                    const u64 f = min((u64)cap, (u64)c * (x.est_deg + 1u));
                    if (dn) {
                        sm.gstart[ng++] = s;
                        open = false;
                    } else if (!open || acc + f > x.budget_sectors || s - sm.gstart[ng - 1] >= WARP) { // (phase B: <= 32 slots per group)
                        sm.gstart[ng++] = s;
                        open = true;
                        acc = f;
                    } else {
                        acc += f;
                    }
                }
                sm.gstart[ng] = a.slots;
                sm.ng = ng;
            }
        }
        __syncthreads();
        const u32 nf = sm.fbase[a.slots];
        if (nf == 0) break;
        if (level >= a.max_levels) { // never reached in practice (2^20 levels): report instead of dropping the frontier silently
            if (blockIdx.x == 0 && threadIdx.x == 0 && a.err) *a.err = 1;
            if (threadIdx.x < (u32)a.slots) sm.prevcnt[threadIdx.x] = 0; // the counts just read belong to a level that did not run: not logged
            __syncthreads();
            break;
        }
        const u64* cur = (level & 1) ? a.front1 : a.front0;
        u64* nxt = (level & 1) ? a.front0 : a.front1;
        u32* nxt_count = ctl->fcount[(level + 1) % 3];
        const int ng = sm.ng;
        if (blockIdx.x == 0) {
            if (threadIdx.x < (u32)a.slots) {
                ctl->fcount[(level + 2) % 3][threadIdx.x] = 0;
                x.c->hubcnt[(level + 2) % 3][threadIdx.x] = 0;
            }

            if (threadIdx.x == 0) {
                ctl->levels_run = level + 1;
                if (a.trace && level < a.trace_cap) {
                    u64 tt;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
                    a.trace[4 * level] = tt;
                    a.trace[4 * level + 1] = nf;
                    u32 nd = 0;
                    for (int s = 0; s < a.slots; ++s) nd += sm.dense[s];
                    a.trace[4 * level + 2] = (u64)ng | ((u64)nd << 32);
                }
            }
        }
        // pipeline over the groups: interval j does phase A of group j and phase B of group j - 1; a dense group is scanned right
        // behind its phase B (one more barrier: its vector is still in the L2, and nothing else competes for it)
        for (int j = 0; j <= ng; ++j) {
            if (threadIdx.x == 0) sm.qdone = 0;
            if (blockIdx.x == 0 && threadIdx.x < P3_NQ + P3_NQH) { // consumed in interval t - 1, used again in t + 2
                if (threadIdx.x < P3_NQ) x.c->cursor[(t + 2) % 3][threadIdx.x] = 0;
                else x.c->hcursor[(t + 2) % 3][threadIdx.x - P3_NQ] = 0;
            }
            __syncthreads();
            // phase A of group j runs on the first P3_AWARPS warps of every CTA while the others already scatter; alone in
            // its interval (first group of a level) it takes the whole CTA
            if (j < ng) {
                const u32 nth = j > 0 ? (u32)P3_AWARPS * WARP : (u32)P3_THREADS;
                if (threadIdx.x < nth) p3_phase_a<OffT>(a, g, x, sm, cur, sm.gstart[j], sm.gstart[j + 1], blockIdx.x, gridDim.x, level, threadIdx.x, nth);
            }
            if (j >= 1) {
                const int s0 = sm.gstart[j - 1], s1 = sm.gstart[j];
                if (sm.dense[s0]) {
                    // (development trace, CTA 0: where a dense interval's time goes -- own phase B, wait, own scan, wait)
                    const bool tr = a.trace && blockIdx.x == 0 && level < 1024;
                    u64 t0 = 0, t1 = 0, t2 = 0, t3 = 0;
                    if (tr) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)); }
                    p3_phase_b<OffT, true>(a, g, x, sm, cur, s0, s1, t % 3, level, nxt, nxt_count);
                    if (tr) { __syncthreads(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); }
                    grid.sync();
                    if (tr) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2)); }
                    p3_scan<OffT>(a, g, x, sm, s0, level, blockIdx.x, gridDim.x, nxt, nxt_count);
                    if (tr) {
                        __syncthreads();
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t3));
                        if (threadIdx.x == 0) {
                            u64* tx = a.trace + 4 * 1024 + 4 * level;
                            tx[0] += t1 - t0; tx[1] += t2 - t1; tx[2] += t3 - t2; tx[3] = t3;
                        }
                    }
                } else {
                    p3_phase_b<OffT, false>(a, g, x, sm, cur, s0, s1, t % 3, level, nxt, nxt_count);
                }
            }
            grid.sync();
            ++t;
            if (j == 0 && blockIdx.x == 0 && threadIdx.x == 0 && a.trace && level < a.trace_cap) {
                u64 tt;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
                a.trace[4 * level + 3] = tt;
            }
        }
        // flush this CTA's per-slot counters of the level
        if (threadIdx.x < (u32)a.slots) {
            const int sl = threadIdx.x;
            const u32 v = sm.cnt_verts[sl];
            if (v) { // (a scan counts the vertices of the NEXT level's frontier; the totals of a launch come out the same)
                atomicAdd(&a.edges[sl], (u64)sm.cnt_edges[sl]);
                atomicAdd(&a.vertices[sl], (u64)v);
                sm.cnt_edges[sl] = 0;
                sm.cnt_verts[sl] = 0;
            }
            if (blockIdx.x == 0 && sm.inlevel[sl]) { // levels in which the slot pushed something
                const int lvl = (int)(a.level_base + level + 1);
                if (a.lastlvl[sl] < lvl) { a.lastlvl[sl] = lvl; atomicAdd(&a.levels[sl], 1ull); }
            }
        }
        __syncthreads();
    }
    // every CTA holds the same log positions; CTA 0 publishes them for the next launch of the wave / the apply pass
    if (a.log_v && blockIdx.x == 0 && threadIdx.x < (u32)a.slots) a.log_cur[threadIdx.x] = sm.logbase[threadIdx.x] + sm.prevcnt[threadIdx.x];
}

} // namespace fora
