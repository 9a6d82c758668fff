// fora_b200/csrc/engine.cu -- C ABI implementation: context, HBM-resident graph, query batching.
//
// Host orchestration only; all arithmetic on vectors happens in the kernels of push.cuh /
// walk.cuh / topk.cuh.  There is no CPU fallback: without a CUDA device fora_ctx_create fails.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <functional>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include "bippr.cuh"
#include "common.cuh"
#include "push.cuh"
#include "push2.cuh"
#include "push3.cuh"
#include "topk.cuh"
#include "walk.cuh"

using namespace fora;

static std::string g_create_error;

// per-slot metadata, mirrored host <-> device in one copy
struct SlotMeta {
    int32_t source[MAX_SLOTS];
    u32 qid[MAX_SLOTS];
    double rmax[MAX_SLOTS];
    int32_t state[MAX_SLOTS];  // 0 unused, 1 active, 2 source without out-edges
    int32_t active[MAX_SLOTS]; // takes part in the current push round
    u64 edges[MAX_SLOTS], vertices[MAX_SLOTS], levels[MAX_SLOTS];
    int32_t lastlvl[MAX_SLOTS];
    double rsum[MAX_SLOTS];
    u64 nnz[MAX_SLOTS], nsrc[MAX_SLOTS], nwalk[MAX_SLOTS], hops[MAX_SLOTS], idx_hits[MAX_SLOTS];
    double next_rmax[MAX_SLOTS]; // speculative seeding threshold of the following round (0: none)
    u32 seed_count[MAX_SLOTS];   // seeds found for it (segments of front0)
    // push2 / tail kernels (push2.cuh)
    u32 force[MAX_SLOTS];        // the tail kernel refused the slot's level for its edge count: the sub-wave kernel runs it
    u32 left[MAX_SLOTS];         // frontier entries left when the tail kernel returned (0: the slot's round is complete)
    int32_t push_err;
    int32_t pad_;
    u32 sp_count[MAX_SLOTS];     // fora_query_batch_sparse: entries >= threshold found per slot
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    bool owned = true;
    void alias(T* ptr, size_t count) { // view into the hot arena
        if (p && owned) cudaFree(p);
        p = ptr; cap = count; owned = false;
    }
    cudaError_t ensure(size_t count) {
        if (count <= cap) return cudaSuccess;
        if (p && owned) cudaFree(p);
        owned = true;
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e == cudaSuccess) cap = count;
        return e;
    }
    void release() {
        if (p && owned) cudaFree(p);
        p = nullptr;
        cap = 0;
        owned = true;
    }
};

struct fora_ctx {
    int device = 0;
    uint64_t seed = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_stage_ready = nullptr, ev_stage_free = nullptr;
    bool stage_busy = false;
    cudaEvent_t ev[8] = {};
    std::string err;
    int num_sms = 0;
    DeviceGraph g;
    std::vector<int32_t> h_old2new, h_new2old; // host copies of the internal relabelling (empty: identity)
    fora_params p{};
    bool params_set = false;
    int slots = 16;
    int alloc_slots = 0;
    // hot arena: [residue | deg | row offsets (u32) | reserve] in ONE allocation so that an L2 access-policy
    // window can pin the arrays every edge / hop touches at random (B200: 126 MB L2)
    DevBuf<unsigned char> arena;
    int32_t* hot_deg = nullptr;
    u32* hot_ptr32 = nullptr;
    size_t win_push_off = 0, win_push_bytes = 0, win_walk_off = 0, win_walk_bytes = 0;
    size_t l2_persist_max = 0, l2_window_max = 0;
    bool l2_policy = true;
    // dense per-slot state (views into the arena)
    DevBuf<double> reserve, residue;
    DevBuf<u64> front0, front1;
    DevBuf<double> inc;
    DevBuf<u32> eoff;
    DevBuf<u64> block_sum;
    DevBuf<int32_t> log_v;  // reserve credit log of the push (push.cuh)
    DevBuf<double> log_r;
    DevBuf<u32> log_cur;
    size_t log_cap = 0;
    DevBuf<u64> trace; // FORA_PUSH_TRACE=1: per-level trace of the last push launch
    bool trace_on = false;
    DevBuf<PushCtl> ctl;
    DevBuf<SlotMeta> meta;
    SlotMeta* h_meta = nullptr; // pinned + mapped; h_meta_dev is the device-side alias
    SlotMeta* h_meta_dev = nullptr;
    DevBuf<double> part_sum;
    DevBuf<u32> part_nnz;
    int red_blocks = 0;
    // plan / walk
    DevBuf<u32> blk_src;
    DevBuf<u64> blk_walk;
    DevBuf<int32_t> srcs;
    DevBuf<u64> woff;
    DevBuf<double> incs;
    DevBuf<u32> chunk_first;
    DevBuf<double> stage;    // results of a finished wave, copied to the host while the next wave computes
    DevBuf<double> ppr;      // top-k rounds: ppr is rebuilt from reserve every round (query.h:533)
    DevBuf<double> ub, lb;   // non --opt top-k: per-node upper / lower PPR bounds (algo.h:48-49)
    DevBuf<unsigned char> in_topk;
    DevBuf<int32_t> stop_slots;
    DevBuf<double> stop_lowk;
    DevBuf<u32> flags;
    // batched select (topk.cuh): per-vector state, candidate lists, output lists
    DevBuf<SelSlot> sel_st;
    DevBuf<SelResult> sel_res;
    DevBuf<int32_t> sel_slots, sel_ci, sel_on;
    DevBuf<u64> sel_ck;
    DevBuf<double> sel_ov;
    u32 sel_p2 = 0;
    DevBuf<u64> idx_used;    // top-k with index: per-(slot,vertex) cursor into the index (rw_counter, query.h:575-603)
    size_t chunk_cap = 0;
    // index
    DevBuf<u64> idx_off, idx_cnt;
    DevBuf<int32_t> idx_dest;
    bool has_index = false;
    // scratch
    DevBuf<u64> counts;
    // backward push scratch (bippr.cuh)
    DevBuf<double> bwd_res, bwd_rv;
    DevBuf<int32_t> bwd_lists;
    int bwd_blocks = 0;
    u32 bwd_epoch = 1; // next unused stamp value of the backward push's touched lists (bippr.cuh)
    DevBuf<u64> scratch64;
    DevBuf<int32_t> scratch32;
    DevBuf<double> scratchd;
    int push_grid = 0;
    // second-generation push (push2.cuh): sub-waves of push_sub slots, tails per slot
    int push_v = 1, push_sub = 2, push2_grid = 0, push_prefetch = 1;
    size_t persist_now = (size_t)-1, persist_walk = 0; // current persisting carve-out / the one the walk phase wants
    int push_win = 1;          // pin the sub-wave's residue vectors with an access-policy window (persisting L2 lines)
    int push_win_reset = 0;    // cudaCtxResetPersistingL2Cache after the push phase
    size_t push_carve = 0;     // persisting carve-out while push2 runs
    u32 tail_nf = 1024, tail_e = 8192;
    DevBuf<u32> hubbuf;
    DevBuf<u64> front_begs;
    // third-generation push (push3.cuh): lockstep sweep over slot groups
    int push3_grid = 0;
    DevBuf<P3Ctl> p3ctl;
    DevBuf<u64> p3hub;
    size_t p3hub_cap = 0;
    double p3_budget = 1.0; // a group's estimated scatter footprint, in residue vectors
    u32 p3_hub_deg = P3_HUB_DEG, p3_hub_piece = P3_HUB_PIECE;
    double p3_dense = 1.0 / 32;
    // first-generation kernel, dense slot-levels: edge lists (push.cuh push_dense_scan / push_el_adds)
    DevBuf<uint4> el;
    DevBuf<u32> el_ctl; // [2][MAX_SLOTS] counts, [2][MAX_SLOTS] overflow flags
    size_t el_cap = 0;
    // bulk walks (index build / Monte-Carlo / BiPPR) through the chunked walk kernel
    DevBuf<u32> bulk_chunk_first;
    DevBuf<unsigned char> bulk_meta;
    DevBuf<u64> bulk_small;
    // compacted query output (fora_query_batch_sparse)
    DevBuf<int32_t> sp_ids;
    DevBuf<double> sp_vals;
    DevBuf<u32> sp_cnt;
    cudaEvent_t ev_sp[2] = {};
    bool sp_busy[2] = {false, false};
    u64 bulk_walks = 0, bulk_hops = 0;
    double bulk_kernel_ms = 0;
    // per-kernel timing: event pairs recorded around the hot kernels, harvested after the next stream sync
    std::vector<cudaEvent_t> kev_pool;
    std::vector<std::pair<int, int> > kev_pending; // (pool index of start event, kind 0 push / 1 walk)
    size_t kev_used = 0;
    double push_kernel_ms = 0, walk_kernel_ms = 0;
    u64 push_kernel_launches = 0, walk_kernel_launches = 0;
    u32 level_base = 0;
    u64 launches = 0;
    u64 qid_base = 0; // global index of the first query of the next batch call (Philox key); see fora_ctx_set_query_base
    // shared walks (fora_ctx_set_shared_walks): a per-wave virtual walk index
    int shared_walks = 0;
    double shared_cost_scale = 0.2; // --balanced: cost of a walk drawn from the pool relative to a private walk
    DevBuf<u32> sw_cnt;        // [n] walks needed from each vertex = max over the wave's slots
    DevBuf<u64> sw_cnt64, sw_off; // [n] the same as the index arrays the walk kernel reads
    DevBuf<u32> sw_flag, sw_pos;  // [n] vertices with walks, their rank
    DevBuf<int32_t> sw_srcs;   // [n] those vertices, compacted
    DevBuf<u64> sw_woff;       // [n + 1] exclusive prefix of their counts
    DevBuf<int32_t> sw_dest;   // destinations, internal ids
    DevBuf<unsigned char> sw_tmp; // cub scratch
    u64 sw_waves = 0;          // waves served so far (Philox key of a wave's walks)
    u64 sw_built_walks = 0, sw_built_hops = 0; // of the last wave
    // resumable push session (fora_push_begin / fora_push_round)
    int32_t session_source = -1;

    int fail(int code, const std::string& msg) {
        err = msg;
        return code;
    }
};

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return ctx->fail(FORA_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__) + " @" + __FILE__ + \
                                             ":" + std::to_string(__LINE__));                                 \
    } while (0)
#define CKL()                                                                                   \
    do {                                                                                        \
        ctx->launches++;                                                                        \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess)                                                                 \
            return ctx->fail(FORA_ECUDA, std::string("launch: ") + cudaGetErrorString(e__) + " @" + __FILE__ + \
                                             ":" + std::to_string(__LINE__));                   \
    } while (0)

static int relabel_graph(fora_ctx* ctx);
static int set_persist_limit(fora_ctx* ctx, size_t bytes);
static int pack_columns(fora_ctx* ctx);
static int permute_csr(fora_ctx* ctx, int32_t n, int64_t ne, const int32_t* src_of, const int32_t* map, const int64_t* in_ptr,
                       const int32_t* in_col, int64_t** out_ptr, int32_t** out_col);

// --balanced cost model, calibrated on B200 with the LiveJournal-shape graph (profiles/): a walk costs
// 1.5 ms / 24 M = 6.5e-11 s (no-zero-hop, ~4.9 hops, relabelled graph, 32 slots); push ~3.5e-11 s per edge in the scatter
// phase plus ~1e-10 s per vertex in the gather phase and ~10 us of barrier latency per level.  With these the loop stops
// where push time ~ walk time (2.0 ms vs 2.1 ms per query measured), which is what query.h:867-877 aims for.
static const double DEFAULT_COST_WALK = 6.5e-11;   // s per online walk
static const double DEFAULT_COST_EDGE = 3.5e-11;   // s per pushed edge
static const double DEFAULT_COST_VERTEX = 1.0e-10; // s per pushed vertex
static const double DEFAULT_COST_LEVEL = 1.0e-5;   // s per frontier level

// =============================================================================================
// context
// =============================================================================================
extern "C" int fora_ctx_create(int device, uint64_t seed, fora_ctx** out) {
    if (!out) return FORA_EINVAL;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)";
        return FORA_ECUDA;
    }
    if (device < 0 || device >= count) {
        g_create_error = "bad device ordinal";
        return FORA_EINVAL;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e);
        return FORA_ECUDA;
    }
    fora_ctx* ctx = new fora_ctx();
    ctx->device = device;
    ctx->seed = seed;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    ctx->num_sms = prop.multiProcessorCount;
    ctx->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
    ctx->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    ctx->l2_policy = getenv("FORA_NO_L2_POLICY") == nullptr && ctx->l2_persist_max > 0;
    // (cudaLimitMaxL2FetchGranularity 32 / 64 / 128 was measured to have no effect on B200 for this engine's random 4- and
    // 8-byte accesses -- profiles/r1_ubench_l2_fetch_granularity.txt -- so it is left at the driver's default.)
    if (!prop.cooperativeLaunch) {
        g_create_error = "device lacks cooperative launch";
        delete ctx;
        return FORA_ECUDA;
    }
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaHostAlloc((void**)&ctx->h_meta, sizeof(SlotMeta), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&ctx->h_meta_dev, ctx->h_meta, 0) != cudaSuccess) {
        g_create_error = "stream / pinned allocation failed";
        delete ctx;
        return FORA_ECUDA;
    }
    ctx->stream = ctx->own_stream;
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->ev_stage_ready, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_stage_free, cudaEventDisableTiming);
    for (auto& ev : ctx->ev) cudaEventCreate(&ev);
    memset(ctx->h_meta, 0, sizeof(SlotMeta));
    ctx->p.alpha = 0.2;
    *out = ctx;
    return FORA_OK;
}

static void free_graph(DeviceGraph& g) {
    cudaFree(g.out_ptr64); cudaFree(g.out_ptr32); cudaFree(g.out_col); cudaFree(g.deg); cudaFree(g.out_colx);
    cudaFree(g.in_ptr64); cudaFree(g.in_ptr32); cudaFree(g.in_col);
    cudaFree(g.old2new); cudaFree(g.new2old);
    g = DeviceGraph();
}

extern "C" void fora_ctx_destroy(fora_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    free_graph(ctx->g);
    ctx->reserve.release(); ctx->residue.release(); ctx->arena.release(); ctx->front0.release(); ctx->front1.release();
    ctx->log_v.release(); ctx->log_r.release(); ctx->log_cur.release();
    ctx->sp_ids.release(); ctx->sp_vals.release(); ctx->sp_cnt.release();
    for (auto& e : ctx->ev_sp) if (e) cudaEventDestroy(e);
    ctx->hubbuf.release(); ctx->front_begs.release(); ctx->bulk_chunk_first.release(); ctx->bulk_meta.release(); ctx->bulk_small.release(); ctx->inc.release(); ctx->eoff.release(); ctx->block_sum.release(); ctx->trace.release(); ctx->ctl.release(); ctx->meta.release();
    ctx->part_sum.release(); ctx->part_nnz.release(); ctx->blk_src.release(); ctx->blk_walk.release();
    ctx->ppr.release(); ctx->ub.release(); ctx->lb.release(); ctx->in_topk.release(); ctx->stop_slots.release(); ctx->stop_lowk.release(); ctx->flags.release(); ctx->stage.release(); ctx->idx_used.release(); ctx->srcs.release(); ctx->woff.release(); ctx->incs.release(); ctx->chunk_first.release();
    ctx->idx_off.release(); ctx->idx_cnt.release(); ctx->idx_dest.release();
    ctx->sel_st.release(); ctx->sel_res.release(); ctx->sel_slots.release(); ctx->sel_ci.release(); ctx->sel_on.release(); ctx->sel_ck.release(); ctx->sel_ov.release();
    ctx->counts.release(); ctx->bwd_res.release(); ctx->bwd_rv.release(); ctx->bwd_lists.release(); ctx->scratch64.release(); ctx->scratch32.release(); ctx->scratchd.release();
    if (ctx->h_meta) cudaFreeHost(ctx->h_meta);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->kev_pool) cudaEventDestroy(ev);
    if (ctx->ev_stage_ready) cudaEventDestroy(ctx->ev_stage_ready);
    if (ctx->ev_stage_free) cudaEventDestroy(ctx->ev_stage_free);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

extern "C" const char* fora_last_error(fora_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int fora_ctx_set_stream(fora_ctx* ctx, void* s) {
    if (!ctx) return FORA_EINVAL;
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return FORA_OK;
}
extern "C" int fora_ctx_set_slots(fora_ctx* ctx, int slots) {
    if (!ctx || slots < 1 || slots > MAX_SLOTS) return ctx ? ctx->fail(FORA_EINVAL, "slots must be in [1,64]") : FORA_EINVAL;
    ctx->slots = slots;
    return FORA_OK;
}
// Philox streams are keyed by (ctx seed, GLOBAL query index): a caller that shards a query list over GPUs or issues it in
// several batch calls sets the index of the first query of the next call here, so that no two queries of a run share a
// random stream and results do not depend on how the list was cut.  Sticky until changed.
extern "C" int fora_ctx_set_query_base(fora_ctx* ctx, uint64_t first_query_index) {
    if (!ctx) return FORA_EINVAL;
    ctx->qid_base = first_query_index;
    return FORA_OK;
}
extern "C" int fora_ctx_set_shared_walks(fora_ctx* ctx, int on) {
    if (!ctx) return FORA_EINVAL;
    ctx->shared_walks = on ? 1 : 0;
    if (getenv("FORA_SHARED_COST_SCALE")) ctx->shared_cost_scale = atof(getenv("FORA_SHARED_COST_SCALE"));
    return FORA_OK;
}
extern "C" int fora_ctx_sync(fora_ctx* ctx) {
    if (!ctx) return FORA_EINVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return FORA_OK;
}

// =============================================================================================
// graph
// =============================================================================================
__global__ void degree_kernel(int32_t n, const int64_t* __restrict__ ptr, int32_t* __restrict__ deg, u32* __restrict__ ptr32) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v <= n; v += gridDim.x * blockDim.x) {
        if (v < n) deg[v] = (int32_t)(ptr[v + 1] - ptr[v]);
        if (ptr32) ptr32[v] = (u32)ptr[v];
    }
}

// a CSR handed in by the caller is checked once on the device: offsets non-decreasing from 0, ids in [0, n)
__global__ void csr_check_kernel(int32_t n, int64_t ne, const int64_t* __restrict__ ptr, const int32_t* __restrict__ col, int* __restrict__ bad) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int64_t v = t0; v < n; v += stride)
        if (ptr[v] > ptr[v + 1] || (v == 0 && ptr[0] != 0)) *bad = 1;
    for (int64_t e = t0; e < ne; e += stride)
        if ((u32)col[e] >= (u32)n) *bad = 2;
}

static int graph_upload_impl(fora_ctx* ctx, int32_t n, int64_t m_decl, const int64_t* out_ptr, const int32_t* out_col,
                             const int64_t* in_ptr, const int32_t* in_col);
extern "C" int fora_graph_upload(fora_ctx* ctx, int32_t n, int64_t m_decl, const int64_t* out_ptr, const int32_t* out_col,
                                 const int64_t* in_ptr, const int32_t* in_col) {
    if (!ctx) return FORA_EINVAL;
    const int rc = graph_upload_impl(ctx, n, m_decl, out_ptr, out_col, in_ptr, in_col);
    if (rc) { // never leave a half-initialised graph behind
        cudaStreamSynchronize(ctx->stream);
        free_graph(ctx->g);
    }
    return rc;
}
static int graph_upload_impl(fora_ctx* ctx, int32_t n, int64_t m_decl, const int64_t* out_ptr, const int32_t* out_col,
                             const int64_t* in_ptr, const int32_t* in_col) {
    if (n <= 0 || !out_ptr || !out_col) return ctx->fail(FORA_EINVAL, "graph_upload: n>0 and out-CSR required");
    if (out_ptr[0] != 0 || out_ptr[n] < 0) return ctx->fail(FORA_EINVAL, "graph_upload: out_ptr must start at 0 and be non-negative");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    free_graph(ctx->g);
    DeviceGraph& g = ctx->g;
    g.n = n;
    g.m_decl = m_decl;
    g.n_edges = out_ptr[n];
    g.off32 = g.n_edges < (int64_t)0xffffffffLL;
    const size_t ne = (size_t)g.n_edges;
    CK(cudaMalloc((void**)&g.out_ptr64, sizeof(int64_t) * (size_t)(n + 1)));
    CK(cudaMalloc((void**)&g.out_col, sizeof(int32_t) * std::max<size_t>(ne, 1)));
    CK(cudaMalloc((void**)&g.deg, sizeof(int32_t) * (size_t)n));
    if (g.off32) CK(cudaMalloc((void**)&g.out_ptr32, sizeof(u32) * (size_t)(n + 1)));
    CK(cudaMemcpyAsync(g.out_ptr64, out_ptr, sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(g.out_col, out_col, sizeof(int32_t) * ne, cudaMemcpyHostToDevice, ctx->stream));
    DevBuf<int> bad;
    CK(bad.ensure(1));
    CK(cudaMemsetAsync(bad.p, 0, sizeof(int), ctx->stream));
    csr_check_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(n, g.n_edges, g.out_ptr64, g.out_col, bad.p);
    CKL();
    if (in_ptr && in_col) {
        if (in_ptr[0] != 0 || in_ptr[n] != g.n_edges) { bad.release(); return ctx->fail(FORA_EINVAL, "graph_upload: in-CSR edge count differs from out-CSR"); }
        CK(cudaMalloc((void**)&g.in_ptr64, sizeof(int64_t) * (size_t)(n + 1)));
        CK(cudaMalloc((void**)&g.in_col, sizeof(int32_t) * std::max<size_t>(ne, 1)));
        if (g.off32) CK(cudaMalloc((void**)&g.in_ptr32, sizeof(u32) * (size_t)(n + 1)));
        CK(cudaMemcpyAsync(g.in_ptr64, in_ptr, sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(g.in_col, in_col, sizeof(int32_t) * ne, cudaMemcpyHostToDevice, ctx->stream));
        csr_check_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(n, g.n_edges, g.in_ptr64, g.in_col, bad.p);
        CKL();
        g.has_in = true;
    }
    int hbad = 0;
    CK(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    bad.release();
    if (hbad) return ctx->fail(FORA_EINVAL, hbad == 1 ? "graph_upload: row offsets are not non-decreasing" : "graph_upload: a column id lies outside [0, n)");
    degree_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(n, g.out_ptr64, g.deg, g.out_ptr32);
    CKL();
    if (g.has_in && g.off32) {
        DevBuf<int32_t> tmp;
        CK(tmp.ensure((size_t)n));
        degree_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(n, g.in_ptr64, tmp.p, g.in_ptr32);
        CKL();
        CK(cudaStreamSynchronize(ctx->stream));
        tmp.release();
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->alloc_slots = 0; // dense state is sized by n
    ctx->has_index = false;
    ctx->session_source = -1;
    ctx->bwd_blocks = 0;
    int rrc = relabel_graph(ctx);
    return rrc ? rrc : pack_columns(ctx);
}

extern "C" int fora_graph_download_csr(fora_ctx* ctx, int64_t* out_ptr, int32_t* out_col, int64_t* in_ptr, int32_t* in_col) {
    if (!ctx || !ctx->g.n) return ctx ? ctx->fail(FORA_EINVAL, "no graph") : FORA_EINVAL;
    CK(cudaSetDevice(ctx->device));
    const DeviceGraph& g = ctx->g;
    const size_t ne = (size_t)g.n_edges;
    if ((in_ptr || in_col) && !g.has_in) return ctx->fail(FORA_EINVAL, "in-CSR was not uploaded");
    // the kernels work on a relabelled copy; the download undoes the permutation (bit-exact original lists)
    int64_t *optr = g.out_ptr64, *iptr = g.in_ptr64, *t_optr = nullptr, *t_iptr = nullptr;
    int32_t *ocol = g.out_col, *icol = g.in_col, *t_ocol = nullptr, *t_icol = nullptr;
    int rc;
    if (g.relabeled) {
        if ((rc = permute_csr(ctx, g.n, g.n_edges, g.old2new, g.new2old, g.out_ptr64, g.out_col, &t_optr, &t_ocol))) return rc;
        optr = t_optr; ocol = t_ocol;
        if (in_ptr || in_col) {
            if ((rc = permute_csr(ctx, g.n, g.n_edges, g.old2new, g.new2old, g.in_ptr64, g.in_col, &t_iptr, &t_icol))) return rc;
            iptr = t_iptr; icol = t_icol;
        }
    }
    if (out_ptr) CK(cudaMemcpyAsync(out_ptr, optr, sizeof(int64_t) * (size_t)(g.n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_col) CK(cudaMemcpyAsync(out_col, ocol, sizeof(int32_t) * ne, cudaMemcpyDeviceToHost, ctx->stream));
    if (in_ptr) CK(cudaMemcpyAsync(in_ptr, iptr, sizeof(int64_t) * (size_t)(g.n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if (in_col) CK(cudaMemcpyAsync(in_col, icol, sizeof(int32_t) * ne, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(t_optr); cudaFree(t_ocol); cudaFree(t_iptr); cudaFree(t_icol);
    return FORA_OK;
}
extern "C" int64_t fora_graph_num_edges(fora_ctx* ctx) { return ctx ? ctx->g.n_edges : FORA_EINVAL; }


// =============================================================================================
// K0: CSR construction on the device from an edge list in file order (Graph::init_graph, graph.h:152-160):
// drop self loops, keep duplicates, keep file order inside every adjacency list.  "File order inside a list" is
// exactly a STABLE sort of the edges by source (out-CSR) / by target (in-CSR), so the build is: flag + stable
// compaction, stable LSD radix sort of (key, value) pairs, histogram + exclusive scan for the row offsets.
// The sort / scan / select primitives are CUB's (the CUDA toolkit's library, like calling cuBLAS for a GEMM):
// this is load-time plumbing, not the query hot path.
// =============================================================================================
__global__ void k0_flag_kernel(int64_t ne, const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int32_t n,
                               unsigned char* __restrict__ keep, int* __restrict__ bad) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
        const int32_t s = src[e], d = dst[e];
        if ((u32)s >= (u32)n || (u32)d >= (u32)n) *bad = 1; // graph.h:155-156
        keep[e] = s != d;                                    // graph.h:157
    }
}
__global__ void k0_count_kernel(int64_t ne, const int32_t* __restrict__ key, unsigned long long* __restrict__ cnt) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&cnt[key[e]], 1ull);
}

static int k0_one_direction(fora_ctx* ctx, int32_t n, int64_t kept, const int32_t* d_key, const int32_t* d_val, int64_t* d_ptr,
                            int32_t* d_col, int key_bits) {
    DevBuf<int32_t> key_out;
    DevBuf<unsigned long long> cnt;
    DevBuf<unsigned char> tmp;
    CK(key_out.ensure((size_t)std::max<int64_t>(kept, 1)));
    CK(cnt.ensure((size_t)n + 1));
    CK(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long) * ((size_t)n + 1), ctx->stream));
    if (kept > 0) {
        k0_count_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(kept, d_key, cnt.p);
        CKL();
        size_t bytes = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, d_key, key_out.p, d_val, d_col, kept, 0, key_bits, ctx->stream));
        CK(tmp.ensure(bytes));
        CK(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, d_key, key_out.p, d_val, d_col, kept, 0, key_bits, ctx->stream)); // stable
    }
    size_t bytes = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt.p, (unsigned long long*)d_ptr, n + 1, ctx->stream));
    CK(tmp.ensure(bytes));
    CK(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, cnt.p, (unsigned long long*)d_ptr, n + 1, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    key_out.release(); cnt.release(); tmp.release();
    return FORA_OK;
}

extern "C" int fora_graph_build_from_edges(fora_ctx* ctx, int32_t n, int64_t m_decl, const int32_t* src, const int32_t* dst,
                                           int64_t n_edges, int with_in) {
    if (!ctx) return FORA_EINVAL;
    if (n <= 0 || n_edges < 0 || (n_edges && (!src || !dst))) return ctx->fail(FORA_EINVAL, "graph_build_from_edges: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    const size_t ne = (size_t)n_edges;
    DevBuf<int32_t> d_src, d_dst, k_src, k_dst;
    DevBuf<unsigned char> keep, tmp;
    DevBuf<int> flags;
    DevBuf<long long> nsel; // kept-edge counts of the two compactions (64-bit: graphs beyond 2^31 edges use the int64-offset kernels)
    CK(nsel.ensure(2));
    CK(d_src.ensure(std::max<size_t>(ne, 1))); CK(d_dst.ensure(std::max<size_t>(ne, 1)));
    CK(k_src.ensure(std::max<size_t>(ne, 1))); CK(k_dst.ensure(std::max<size_t>(ne, 1)));
    CK(keep.ensure(std::max<size_t>(ne, 1))); CK(flags.ensure(4));
    CK(cudaMemcpyAsync(d_src.p, src, sizeof(int32_t) * ne, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_dst.p, dst, sizeof(int32_t) * ne, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(flags.p, 0, sizeof(int) * 4, ctx->stream));
    int64_t kept = 0;
    if (ne) {
        k0_flag_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(n_edges, d_src.p, d_dst.p, n, keep.p, flags.p);
        CKL();
        size_t bytes = 0;
        CK(cub::DeviceSelect::Flagged(nullptr, bytes, d_src.p, keep.p, k_src.p, nsel.p, n_edges, ctx->stream));
        CK(tmp.ensure(bytes));
        CK(cub::DeviceSelect::Flagged(tmp.p, bytes, d_src.p, keep.p, k_src.p, nsel.p, n_edges, ctx->stream)); // order preserving
        CK(cub::DeviceSelect::Flagged(tmp.p, bytes, d_dst.p, keep.p, k_dst.p, nsel.p + 1, n_edges, ctx->stream));
        int hf[4];
        long long hsel[2];
        CK(cudaMemcpyAsync(hf, flags.p, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(hsel, nsel.p, sizeof(long long) * 2, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (hf[0]) return ctx->fail(FORA_ERANGE, "edge list contains a node id >= n (graph.h:155-156)");
        kept = hsel[0];
    }
    free_graph(ctx->g);
    DeviceGraph& g = ctx->g;
    g.n = n; g.m_decl = m_decl; g.n_edges = kept;
    g.off32 = kept < (int64_t)0xffffffffLL;
    int key_bits = 1;
    while ((1ll << key_bits) < (long long)n) ++key_bits;
    CK(cudaMalloc((void**)&g.out_ptr64, sizeof(int64_t) * ((size_t)n + 1)));
    CK(cudaMalloc((void**)&g.out_col, sizeof(int32_t) * std::max<size_t>((size_t)kept, 1)));
    CK(cudaMalloc((void**)&g.deg, sizeof(int32_t) * (size_t)n));
    if (g.off32) CK(cudaMalloc((void**)&g.out_ptr32, sizeof(u32) * ((size_t)n + 1)));
    int rc = k0_one_direction(ctx, n, kept, k_src.p, k_dst.p, g.out_ptr64, g.out_col, key_bits);
    if (rc) return rc;
    degree_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(n, g.out_ptr64, g.deg, g.out_ptr32);
    CKL();
    if (with_in) {
        CK(cudaMalloc((void**)&g.in_ptr64, sizeof(int64_t) * ((size_t)n + 1)));
        CK(cudaMalloc((void**)&g.in_col, sizeof(int32_t) * std::max<size_t>((size_t)kept, 1)));
        if (g.off32) CK(cudaMalloc((void**)&g.in_ptr32, sizeof(u32) * ((size_t)n + 1)));
        if ((rc = k0_one_direction(ctx, n, kept, k_dst.p, k_src.p, g.in_ptr64, g.in_col, key_bits))) return rc;
        if (g.off32) {
            DevBuf<int32_t> t2;
            CK(t2.ensure((size_t)n));
            degree_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(n, g.in_ptr64, t2.p, g.in_ptr32);
            CKL();
            CK(cudaStreamSynchronize(ctx->stream));
            t2.release();
        }
        g.has_in = true;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    d_src.release(); d_dst.release(); k_src.release(); k_dst.release(); keep.release(); tmp.release(); flags.release(); nsel.release();
    ctx->alloc_slots = 0;
    ctx->has_index = false;
    ctx->session_source = -1;
    ctx->bwd_blocks = 0;
    int rrc = relabel_graph(ctx);
    return rrc ? rrc : pack_columns(ctx);
}


// =============================================================================================
// Internal relabelling by descending in-degree.  Where a walk or a push lands is distributed like the
// in-degree, so after the renumbering the hot entries of every per-vertex vector (residue, ppr, row offsets) and
// the adjacency lists of the hot vertices are contiguous and stay in the L2 (measured on the LJ-shape graph:
// +11 % queries/s).  Every adjacency list keeps its internal order.  The ABI keeps speaking ORIGINAL ids: sources,
// dense vectors, top-k ids, index files and the CSR download are mapped at the boundary.
// =============================================================================================
__global__ void indeg_kernel(int64_t ne, const int32_t* __restrict__ col, u32* __restrict__ indeg) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) atomicAdd(&indeg[col[e]], 1u);
}
// mode 0: descending in-degree.  mode 1: descending in-degree / out-degree -- a walker sits at v about as often as v's
// in-degree says and then reads ONE of its d_out neighbour slots, so this ratio is the reference rate of every 4-byte
// slot of v's adjacency list: sorting by it packs the hot part of the column array (not only of the per-vertex vectors).
__global__ void relabel_key_kernel(int32_t n, const u32* __restrict__ indeg, const int32_t* __restrict__ outdeg, int mode,
                                   u32* __restrict__ keys, int32_t* __restrict__ ids) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
        // ascending key == descending score; the stable sort keeps id order among ties
        if (mode == 1) keys[v] = 0xffffffffu - __float_as_uint((float)indeg[v] / (float)max(outdeg[v], 1)); // non-negative floats order like their bits
        else keys[v] = 0xffffffffu - indeg[v];
        ids[v] = v;
    }
}
__global__ void invert_perm_kernel(int32_t n, const int32_t* __restrict__ fwd, int32_t* __restrict__ inv) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) inv[fwd[v]] = v;
}
// output vertex i takes the adjacency list of input vertex src_of[i]; every neighbour id goes through map[]
__global__ void permute_deg_kernel(int32_t n, const int32_t* __restrict__ src_of, const int64_t* __restrict__ in_ptr,
                                   unsigned long long* __restrict__ cnt) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x)
        cnt[i] = i < n ? (unsigned long long)(in_ptr[src_of[i] + 1] - in_ptr[src_of[i]]) : 0ull;
}
__global__ void permute_fill_kernel(int32_t n, const int32_t* __restrict__ src_of, const int32_t* __restrict__ map,
                                    const int64_t* __restrict__ in_ptr, const int32_t* __restrict__ in_col,
                                    const int64_t* __restrict__ out_ptr, int32_t* __restrict__ out_col) {
    const int lane = threadIdx.x & 31;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += (gridDim.x * blockDim.x) >> 5) {
        const int64_t ib = in_ptr[src_of[i]], d = in_ptr[src_of[i] + 1] - ib, ob = out_ptr[i];
        for (int64_t k = lane; k < d; k += 32) out_col[ob + k] = map[in_col[ib + k]];
    }
}
template <typename T>
__global__ void gather_kernel(size_t n, const T* __restrict__ src, const int32_t* __restrict__ idx, T* __restrict__ dst) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}
__global__ void map_ids_kernel(size_t n, const int32_t* __restrict__ in, const int32_t* __restrict__ map, int32_t* __restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i] < 0 ? in[i] : map[in[i]];
}

// CSR (in_ptr, in_col) -> freshly allocated (out_ptr, out_col) under the permutation (src_of, map)
static int permute_csr(fora_ctx* ctx, int32_t n, int64_t ne, const int32_t* src_of, const int32_t* map, const int64_t* in_ptr,
                       const int32_t* in_col, int64_t** out_ptr, int32_t** out_col) {
    DevBuf<unsigned long long> cnt;
    DevBuf<unsigned char> tmp;
    CK(cnt.ensure((size_t)n + 1));
    CK(cudaMalloc((void**)out_ptr, sizeof(int64_t) * ((size_t)n + 1)));
    CK(cudaMalloc((void**)out_col, sizeof(int32_t) * std::max<size_t>((size_t)ne, 1)));
    permute_deg_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(n, src_of, in_ptr, cnt.p);
    CKL();
    size_t bytes = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt.p, (unsigned long long*)*out_ptr, n + 1, ctx->stream));
    CK(tmp.ensure(bytes));
    CK(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, cnt.p, (unsigned long long*)*out_ptr, n + 1, ctx->stream));
    permute_fill_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(n, src_of, map, in_ptr, in_col, *out_ptr, *out_col);
    CKL();
    CK(cudaStreamSynchronize(ctx->stream));
    cnt.release(); tmp.release();
    return FORA_OK;
}

static int relabel_graph(fora_ctx* ctx) {
    DeviceGraph& g = ctx->g;
    ctx->h_old2new.clear();
    ctx->h_new2old.clear();
    if (getenv("FORA_NO_RELABEL") || g.n < 2) return FORA_OK;
    const int32_t n = g.n;
    DevBuf<u32> indeg, keys, keys_out;
    DevBuf<int32_t> ids;
    DevBuf<unsigned char> tmp;
    CK(indeg.ensure((size_t)n)); CK(keys.ensure((size_t)n)); CK(keys_out.ensure((size_t)n)); CK(ids.ensure((size_t)n));
    CK(cudaMalloc((void**)&g.old2new, sizeof(int32_t) * (size_t)n));
    CK(cudaMalloc((void**)&g.new2old, sizeof(int32_t) * (size_t)n));
    CK(cudaMemsetAsync(indeg.p, 0, sizeof(u32) * (size_t)n, ctx->stream));
    if (g.n_edges > 0) {
        indeg_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(g.n_edges, g.out_col, indeg.p);
        CKL();
    }
    const int key_mode = getenv("FORA_RELABEL_KEY") ? atoi(getenv("FORA_RELABEL_KEY")) : 1; // measured: +2 % on push and on walks vs plain in-degree
    relabel_key_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(n, indeg.p, g.deg, key_mode, keys.p, ids.p);
    CKL();
    size_t bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys.p, keys_out.p, ids.p, g.new2old, n, 0, 32, ctx->stream));
    CK(tmp.ensure(bytes));
    CK(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys.p, keys_out.p, ids.p, g.new2old, n, 0, 32, ctx->stream));
    invert_perm_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(n, g.new2old, g.old2new);
    CKL();
    int64_t* nptr = nullptr;
    int32_t* ncol = nullptr;
    int rc = permute_csr(ctx, n, g.n_edges, g.new2old, g.old2new, g.out_ptr64, g.out_col, &nptr, &ncol);
    if (rc) return rc;
    cudaFree(g.out_ptr64); cudaFree(g.out_col);
    g.out_ptr64 = nptr; g.out_col = ncol;
    degree_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(n, g.out_ptr64, g.deg, g.out_ptr32);
    CKL();
    if (g.has_in) {
        if ((rc = permute_csr(ctx, n, g.n_edges, g.new2old, g.old2new, g.in_ptr64, g.in_col, &nptr, &ncol))) return rc;
        cudaFree(g.in_ptr64); cudaFree(g.in_col);
        g.in_ptr64 = nptr; g.in_col = ncol;
        if (g.off32) {
            DevBuf<int32_t> t2;
            CK(t2.ensure((size_t)n));
            degree_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(n, g.in_ptr64, t2.p, g.in_ptr32);
            CKL();
            CK(cudaStreamSynchronize(ctx->stream));
            t2.release();
        }
    }
    ctx->h_old2new.resize((size_t)n);
    ctx->h_new2old.resize((size_t)n);
    CK(cudaMemcpyAsync(ctx->h_old2new.data(), g.old2new, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_new2old.data(), g.new2old, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    g.relabeled = true;
    indeg.release(); keys.release(); keys_out.release(); ids.release(); tmp.release();
    return FORA_OK;
}

// Packed columns for the push kernel: a scatter needs the target's out-degree for the threshold test, which is a
// second random access per edge.  Vertex ids need only ceil(log2 n) bits, so the degree (saturated at dmax, meaning
// "look it up") rides in the spare high bits of a copy of the column array that only the push kernel reads.
__global__ void pack_columns_kernel(int64_t ne, const int32_t* __restrict__ col, const int32_t* __restrict__ deg, u32 shift,
                                    int32_t* __restrict__ colx) {
    const u32 dmax = 0xffffffffu >> shift;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
        const u32 u = (u32)col[e];
        colx[e] = (int32_t)(u | (min((u32)deg[u], dmax) << shift));
    }
}
static int pack_columns(fora_ctx* ctx) {
    DeviceGraph& g = ctx->g;
    g.deg_shift = 0;
    if ((getenv("FORA_PUSH_PACK") && atoi(getenv("FORA_PUSH_PACK")) == 0) || g.n_edges == 0) return FORA_OK; // on by default (+6 % push edges/s)
    u32 bits = 1;
    while ((1ull << bits) < (unsigned long long)g.n) ++bits;
    // test hook: FORA_PACK_SHIFT forces more id bits (fewer degree bits), so that small graphs exercise the saturated path
    if (getenv("FORA_PACK_SHIFT")) bits = std::max(bits, std::min(31u, (u32)atoi(getenv("FORA_PACK_SHIFT"))));
    else if (32 - bits < 4) return FORA_OK; // fewer than 4 spare bits: most degrees would saturate
    CK(cudaMalloc((void**)&g.out_colx, sizeof(int32_t) * (size_t)g.n_edges));
    pack_columns_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(g.n_edges, g.out_col, g.deg, bits, g.out_colx);
    CKL();
    CK(cudaStreamSynchronize(ctx->stream));
    g.deg_shift = bits;
    return FORA_OK;
}

static inline int32_t to_internal(const fora_ctx* ctx, int32_t v) { return ctx->g.relabeled ? ctx->h_old2new[(size_t)v] : v; }
static inline int32_t to_original(const fora_ctx* ctx, int32_t v) { return ctx->g.relabeled ? ctx->h_new2old[(size_t)v] : v; }

// dense per-vertex vector, internal order (device) -> original order (device); dst != src
template <typename T>
static int vec_to_original(fora_ctx* ctx, const T* d_src, T* d_dst, size_t n) {
    if (!ctx->g.relabeled) {
        CK(cudaMemcpyAsync(d_dst, d_src, sizeof(T) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        return FORA_OK;
    }
    gather_kernel<T><<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(n, d_src, ctx->g.old2new, d_dst); // dst[old] = src[old2new[old]]
    CKL();
    return FORA_OK;
}
template <typename T>
static int vec_to_internal(fora_ctx* ctx, const T* d_src, T* d_dst, size_t n) {
    if (!ctx->g.relabeled) {
        CK(cudaMemcpyAsync(d_dst, d_src, sizeof(T) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        return FORA_OK;
    }
    gather_kernel<T><<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(n, d_src, ctx->g.new2old, d_dst); // dst[new] = src[new2old[new]]
    CKL();
    return FORA_OK;
}

// =============================================================================================
// parameters
// =============================================================================================
extern "C" int fora_params_set(fora_ctx* ctx, const fora_params* p) {
    if (!ctx || !p) return FORA_EINVAL;
    if (!(p->alpha > 0 && p->alpha < 1)) return ctx->fail(FORA_EINVAL, "alpha must be in (0,1)");
    ctx->p = *p;
    if (ctx->p.cost_walk == 0 && ctx->p.cost_edge == 0 && ctx->p.cost_vertex == 0 && ctx->p.cost_level == 0) {
        ctx->p.cost_walk = DEFAULT_COST_WALK;
        ctx->p.cost_edge = DEFAULT_COST_EDGE;
        ctx->p.cost_vertex = DEFAULT_COST_VERTEX;
        ctx->p.cost_level = DEFAULT_COST_LEVEL;
    }
    // development knobs: FORA_COST_WALK / _EDGE / _VERTEX / _LEVEL override the calibration
    if (getenv("FORA_COST_WALK")) ctx->p.cost_walk = atof(getenv("FORA_COST_WALK"));
    if (getenv("FORA_COST_EDGE")) ctx->p.cost_edge = atof(getenv("FORA_COST_EDGE"));
    if (getenv("FORA_COST_VERTEX")) ctx->p.cost_vertex = atof(getenv("FORA_COST_VERTEX"));
    if (getenv("FORA_COST_LEVEL")) ctx->p.cost_level = atof(getenv("FORA_COST_LEVEL"));
    ctx->params_set = true;
    return FORA_OK;
}
extern "C" int fora_params_get(fora_ctx* ctx, fora_params* p) {
    if (!ctx || !p) return FORA_EINVAL;
    *p = ctx->p;
    return FORA_OK;
}

// =============================================================================================
// slot state
// =============================================================================================
static int ensure_slots(fora_ctx* ctx, double omega_max) {
    const DeviceGraph& g = ctx->g;
    if (!g.n) return ctx->fail(FORA_EINVAL, "no graph uploaded");
    const int S = ctx->slots;
    const size_t n = (size_t)g.n;
    if (ctx->alloc_slots < S) {
        const size_t fcap = n * S;
        if (fcap >= 0xffffffffull) return ctx->fail(FORA_EINVAL, "slots*n must stay below 2^32");
        {
            auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
            const size_t b_res = pad(n * S * sizeof(double)), b_deg = pad(n * sizeof(int32_t)), b_ptr = pad((n + 1) * sizeof(u32));
            ctx->reserve.release(); ctx->residue.release();
            CK(ctx->arena.ensure(2 * b_res + b_deg + b_ptr));
            unsigned char* base = ctx->arena.p;
            ctx->residue.alias((double*)base, n * S);
            ctx->hot_deg = (int32_t*)(base + b_res);
            ctx->hot_ptr32 = (u32*)(base + b_res + b_deg);
            ctx->reserve.alias((double*)(base + b_res + b_deg + b_ptr), n * S);
            CK(cudaMemcpyAsync(ctx->hot_deg, g.deg, n * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
            if (g.off32) CK(cudaMemcpyAsync(ctx->hot_ptr32, g.out_ptr32, (n + 1) * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
            // push touches residue + deg per edge, walks touch row offsets per hop and ppr per walk.  Pin
            // them when they fit the persisting carve-out (one slot at LJ scale); with more slots only the
            // slot-independent arrays (deg, row offsets) are pinned.
            ctx->win_push_off = 0;
            ctx->win_push_bytes = b_res + b_deg;
            ctx->win_walk_off = b_res + b_deg;
            ctx->win_walk_bytes = b_ptr + b_res;
            // The persisting carve-out is taken away from the normal L2 (79 of 126 MB on B200), so it only
            // pays when the pinned arrays fit: one slot at LiveJournal scale.  Otherwise leave the L2 alone
            // (measured: 4 slots 134 q/s without a window vs 90 q/s with deg+offsets pinned).
            const bool fits = ctx->l2_policy && ctx->win_push_bytes <= ctx->l2_persist_max && ctx->win_walk_bytes <= ctx->l2_persist_max;
            if (!fits) ctx->win_push_bytes = ctx->win_walk_bytes = 0;
            size_t limit = fits ? ctx->l2_persist_max : 0;
            if (!fits && ctx->l2_policy && !getenv("FORA_NO_WALK_PIN")) {
                // many slots: pin only the row offsets (19 MB at LJ scale, read on every hop) during the walks; a small
                // carve-out costs the push kernel nothing measurable (+3.5 % walk steps/s)
                ctx->win_walk_off = b_res + b_deg;
                ctx->win_walk_bytes = b_ptr;
                limit = std::min(ctx->l2_persist_max, b_ptr + ((size_t)1 << 20));
                if (ctx->win_walk_bytes > limit) { ctx->win_walk_bytes = 0; limit = 0; }
            }
            ctx->push_v = getenv("FORA_PUSH_V") ? atoi(getenv("FORA_PUSH_V")) : 1; // 2: sub-waves + tails (push2.cuh), still being tuned
            ctx->push_win = getenv("FORA_PUSH_WIN") ? atoi(getenv("FORA_PUSH_WIN")) : 1;
            ctx->push_win_reset = getenv("FORA_PUSH_WIN_RESET") ? atoi(getenv("FORA_PUSH_WIN_RESET")) : 0;
            if (ctx->push_v == 2 && ctx->push_win && ctx->l2_policy && !fits) {
                // the sub-wave kernel keeps k residue vectors (k x 8n bytes) as persisting lines: without the window the streamed
                // columns / frontier / log traffic of a level (~2x the vectors' size) evicts them (L2 hit rate 45 %, r2b ncu)
                ctx->push_carve = getenv("FORA_PUSH_CARVE_MB") ? (size_t)atol(getenv("FORA_PUSH_CARVE_MB")) << 20 : ctx->l2_persist_max;
                ctx->push_carve = std::min(ctx->push_carve, ctx->l2_persist_max);
            }
            ctx->persist_walk = limit;
            if (fits) { ctx->push_carve = 0; limit = ctx->l2_persist_max; ctx->persist_walk = limit; }
            int lrc = set_persist_limit(ctx, ctx->persist_walk);
            if (lrc) return lrc;
        }
        CK(ctx->front0.ensure(fcap));
        CK(ctx->front1.ensure(fcap));
        CK(ctx->inc.ensure(fcap));
        CK(ctx->eoff.ensure(fcap + 1));
        CK(ctx->block_sum.ensure(MAX_PUSH_CTAS));
        // reserve credit log: room for 2n pushes per slot and wave (LJ-shape --balanced queries push ~1.06 n vertices); more spills
        // into the direct update.  FORA_PUSH_LOG=0 disables the log.
        CK(ctx->log_cur.ensure(MAX_SLOTS));
        CK(cudaMemsetAsync(ctx->log_cur.p, 0, sizeof(u32) * MAX_SLOTS, ctx->stream));
        ctx->log_cap = (getenv("FORA_PUSH_LOG") && atoi(getenv("FORA_PUSH_LOG")) == 0) ? 0 : std::min<size_t>(2 * n, 0xfffffff0u);
        if (ctx->log_cap && getenv("FORA_PUSH_LOG_CAP")) ctx->log_cap = std::max<size_t>(1, (size_t)atoll(getenv("FORA_PUSH_LOG_CAP"))); // test hook: force overflow
        if (ctx->log_cap) {
            CK(ctx->log_v.ensure(ctx->log_cap * S));
            CK(ctx->log_r.ensure(ctx->log_cap * S));
        }
        ctx->el_cap = 0;
        if (getenv("FORA_PUSH_DENSE") && atof(getenv("FORA_PUSH_DENSE")) >= 0 && g.off32 && !(getenv("FORA_PUSH_EL") && atoi(getenv("FORA_PUSH_EL")) == 0)) {
            // a slot-level's edges: at most the graph's; n (16 bytes each) covers every level of the graphs measured (LJ shape peaks
            // at 0.85 n); a level that does not fit takes the tiles
            ctx->el_cap = (size_t)std::min<int64_t>(g.n_edges, (int64_t)(getenv("FORA_PUSH_EL_CAP") ? atof(getenv("FORA_PUSH_EL_CAP")) * (double)n : (double)n));
            CK(ctx->el.ensure(2 * (size_t)S * ctx->el_cap));
            CK(ctx->el_ctl.ensure(4 * MAX_SLOTS));
            CK(cudaMemsetAsync(ctx->el_ctl.p, 0, sizeof(u32) * 4 * MAX_SLOTS, ctx->stream));
        }
        ctx->trace_on = getenv("FORA_PUSH_TRACE") != nullptr;
        if (ctx->trace_on) { CK(ctx->trace.ensure(4 * 4096)); CK(cudaMemset(ctx->trace.p, 0, sizeof(u64) * 4 * 4096)); }
        CK(ctx->ctl.ensure(1));
        CK(ctx->meta.ensure(1));
        ctx->red_blocks = std::max(1, std::min<int>(ctx->num_sms * 4, (int)((n + 4095) / 4096)));
        CK(ctx->part_sum.ensure((size_t)ctx->red_blocks * S));
        CK(ctx->part_nnz.ensure((size_t)ctx->red_blocks * S));
        const size_t nblk = (n + PLAN_THREADS * PLAN_VPT - 1) / (PLAN_THREADS * PLAN_VPT);
        CK(ctx->blk_src.ensure(nblk * S));
        CK(ctx->blk_walk.ensure(nblk * S));
        CK(ctx->srcs.ensure(n * S));
        CK(ctx->woff.ensure((n + 1) * S));
        CK(ctx->incs.ensure(n * S));
        ctx->alloc_slots = S;
        ctx->chunk_cap = 0;
        // occupancy-sized cooperative grid
        int per_sm = 0;
        if (g.off32) {
            CK(cudaFuncSetAttribute(push_kernel<u32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PushSmem<u32>)));
            CK(cudaFuncSetAttribute(push_kernel<u32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PushSmem<u32>)));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, push_kernel<u32, true>, PUSH_THREADS, sizeof(PushSmem<u32>)));
        } else {
            CK(cudaFuncSetAttribute(push_kernel<int64_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PushSmem<int64_t>)));
            CK(cudaFuncSetAttribute(push_kernel<int64_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PushSmem<int64_t>)));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, push_kernel<int64_t, true>, PUSH_THREADS, sizeof(PushSmem<int64_t>)));
        }
        if (per_sm < 1) return ctx->fail(FORA_ECUDA, "push kernel does not fit on an SM");
        ctx->push_grid = std::min(per_sm * ctx->num_sms, MAX_PUSH_CTAS);
        // second-generation push: one 1024-thread CTA per SM for the sub-wave kernel, one CTA per slot for the tails
        ctx->push_sub = std::max(1, std::min(P2_MAX_SUB, getenv("FORA_PUSH_SUB") ? atoi(getenv("FORA_PUSH_SUB")) : 2));
        ctx->push_prefetch = getenv("FORA_PUSH_PREFETCH") ? atoi(getenv("FORA_PUSH_PREFETCH")) : 1;
        ctx->tail_nf = std::min<u32>(TAIL_NF_CAP, getenv("FORA_TAIL_NF") ? (u32)atoi(getenv("FORA_TAIL_NF")) : 1024u);
        ctx->tail_e = getenv("FORA_TAIL_E") ? (u32)atoi(getenv("FORA_TAIL_E")) : 8192u;
        CK(ctx->front_begs.ensure(n * (size_t)std::min(S, P2_MAX_SUB))); // a sub-wave's frontier at most
        CK(ctx->hubbuf.ensure(2 * P2_HUB_CAP + 2));
        CK(cudaMemsetAsync(ctx->hubbuf.p, 0, sizeof(u32) * (2 * P2_HUB_CAP + 2), ctx->stream));
        int per_sm2 = 0;
        if (g.off32) {
            CK(cudaFuncSetAttribute(push2_kernel<u32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P2Smem)));
            CK(cudaFuncSetAttribute(push_tail_kernel<u32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TailSmem)));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, push2_kernel<u32>, P2_THREADS, sizeof(P2Smem)));
        } else {
            CK(cudaFuncSetAttribute(push2_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P2Smem)));
            CK(cudaFuncSetAttribute(push_tail_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TailSmem)));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, push2_kernel<int64_t>, P2_THREADS, sizeof(P2Smem)));
        }
        if (per_sm2 < 1) return ctx->fail(FORA_ECUDA, "push2 kernel does not fit on an SM");
        ctx->push2_grid = ctx->num_sms; // one CTA per SM: the grid barrier has 148 participants
        // third-generation push: hub piece lists (3 rotating sets; a slot-level lists every vertex at most once, so
        // sum ceil(d/1024) over vertices with d > 2048 bounds a slot's pieces)
        {
            if (getenv("FORA_P3_HUB")) { // test hook: "deg,piece" (small values exercise the piece path on small graphs)
                unsigned hd = 0, hp = 0;
                if (sscanf(getenv("FORA_P3_HUB"), "%u,%u", &hd, &hp) == 2 && hd >= 1 && hp >= 1) { ctx->p3_hub_deg = hd; ctx->p3_hub_piece = hp; }
            }
            // exact bound of a slot-level's piece list: every vertex is in a frontier at most once per level
            CK(ctx->scratch64.ensure(1));
            CK(cudaMemsetAsync(ctx->scratch64.p, 0, sizeof(u64), ctx->stream));
            hub_pieces_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(g.n, ctx->hot_deg, ctx->p3_hub_deg, ctx->p3_hub_piece, ctx->scratch64.p);
            CKL();
            u64 pieces = 0;
            CK(cudaMemcpyAsync(&pieces, ctx->scratch64.p, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            ctx->p3hub_cap = std::min<size_t>((size_t)pieces + 16, 0xfffffff0u);
            CK(ctx->p3hub.ensure(2 * (size_t)S * ctx->p3hub_cap));
            CK(ctx->p3ctl.ensure(1));
            ctx->p3_budget = getenv("FORA_P3_BUDGET") ? atof(getenv("FORA_P3_BUDGET")) : 1.0;
            // a slot-level with at least this fraction of the vertices in its frontier runs in dense mode (RED + scan); < 0: never
            ctx->p3_dense = getenv("FORA_P3_DENSE") ? atof(getenv("FORA_P3_DENSE")) : 1.0 / 32;
            int per_sm3 = 0;
            if (g.off32) {
                CK(cudaFuncSetAttribute(push3_kernel<u32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P3Smem)));
                CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm3, push3_kernel<u32>, P3_THREADS, sizeof(P3Smem)));
            } else {
                CK(cudaFuncSetAttribute(push3_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P3Smem)));
                CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm3, push3_kernel<int64_t>, P3_THREADS, sizeof(P3Smem)));
            }
            if (per_sm3 < 1) return ctx->fail(FORA_ECUDA, "push3 kernel does not fit on an SM");
            const int want = getenv("FORA_P3_CTAS") ? atoi(getenv("FORA_P3_CTAS")) : 2;
            ctx->push3_grid = std::min(std::min(per_sm3, std::max(1, want)) * ctx->num_sms, MAX_PUSH_CTAS);
        }
    }
    // walks per slot <= omega*rsum + #sources <= omega + n
    const size_t need = std::max((size_t)WALK_TARGET_CHUNKS + 1, (size_t)((omega_max + (double)n) / WALK_CHUNK)) + 4; // see walk_chunk_size
    if (need > ctx->chunk_cap || ctx->chunk_first.cap < need * S) {
        CK(ctx->chunk_first.ensure(need * S));
        ctx->chunk_cap = need;
    }
    return FORA_OK;
}

// The persisting carve-out is a device-wide limit: the push phase wants it large (hot residue prefixes / a sub-wave's vectors),
// the walk phase small (row offsets only -- what is set aside is lost to the walks' normal traffic).  Switched per phase.
static int set_persist_limit(fora_ctx* ctx, size_t bytes) {
    if (!ctx->l2_persist_max || !ctx->l2_policy) return FORA_OK;
    bytes = std::min(bytes, ctx->l2_persist_max);
    if (bytes == ctx->persist_now) return FORA_OK;
    CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes));
    ctx->persist_now = bytes;
    return FORA_OK;
}
// Pin [arena+off, +bytes) in L2 for the kernels launched next on the work stream (bytes == 0 clears).
static int set_l2_window(fora_ctx* ctx, size_t off, size_t bytes, size_t carve = 0) {
    if (!ctx->l2_policy || bytes == 0) return FORA_OK;
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof av);
    if (bytes) {
        const size_t wb = std::min(bytes, ctx->l2_window_max);
        av.accessPolicyWindow.base_ptr = ctx->arena.p + off;
        av.accessPolicyWindow.num_bytes = wb;
        av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)(carve ? carve : ctx->l2_persist_max) / (double)wb);
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } else {
        av.accessPolicyWindow.num_bytes = 0;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        av.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    }
    CK(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av));
    return FORA_OK;
}


// ---- per-kernel timing -----------------------------------------------------------------------
static cudaEvent_t kev_get(fora_ctx* ctx) {
    if (ctx->kev_used == ctx->kev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ctx->kev_pool.push_back(e);
    }
    return ctx->kev_pool[ctx->kev_used++];
}
static void kev_begin(fora_ctx* ctx, int kind) {
    ctx->kev_pending.push_back(std::make_pair((int)ctx->kev_used, kind));
    cudaEventRecord(kev_get(ctx), ctx->stream);
}
static void kev_end(fora_ctx* ctx) { cudaEventRecord(kev_get(ctx), ctx->stream); }
// call only after the stream has been synchronised
static void kev_harvest(fora_ctx* ctx) {
    for (auto& pr : ctx->kev_pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->kev_pool[pr.first], ctx->kev_pool[pr.first + 1]) == cudaSuccess) {
            if (pr.second == 0) { ctx->push_kernel_ms += ms; ctx->push_kernel_launches++; }
            else if (pr.second == 2) ctx->push_kernel_ms += ms; // further kernels of the same push round
            else { ctx->walk_kernel_ms += ms; ctx->walk_kernel_launches++; }
        }
    }
    ctx->kev_pending.clear();
    ctx->kev_used = 0;
}

static int meta_h2d(fora_ctx* ctx) {
    CK(cudaMemcpyAsync(ctx->meta.p, ctx->h_meta, sizeof(SlotMeta), cudaMemcpyHostToDevice, ctx->stream));
    return FORA_OK;
}
// The per-round control block goes to the host through a kernel that stores into mapped pinned memory, not through
// the device->host copy engine: that engine is a FIFO shared by all streams, and a 4 KB control read queued behind
// the 1.2 GB result copy of the previous wave (copy stream) would stall the next wave for the whole transfer.
__global__ void meta_out_kernel(const u32* __restrict__ src, u32* __restrict__ dst, int words) {
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}
static int meta_d2h_sync(fora_ctx* ctx) {
    static_assert(sizeof(SlotMeta) % sizeof(u32) == 0, "SlotMeta is copied word-wise");
    meta_out_kernel<<<1, 256, 0, ctx->stream>>>((const u32*)ctx->meta.p, (u32*)ctx->h_meta_dev, (int)(sizeof(SlotMeta) / sizeof(u32)));
    CKL();
    CK(cudaStreamSynchronize(ctx->stream));
    kev_harvest(ctx);
    if (ctx->h_meta->push_err) return ctx->fail(FORA_ECUDA, "push: level cap reached with a non-empty frontier");
    return FORA_OK;
}

static PushArgs make_push_args(fora_ctx* ctx) {
    PushArgs a{};
    SlotMeta* m = ctx->meta.p;
    a.n = ctx->g.n;
    a.slots = ctx->slots;
    a.alpha = ctx->p.alpha;
    a.reserve = ctx->reserve.p;
    a.residue = ctx->residue.p;
    a.deg = ctx->hot_deg;
    a.front0 = ctx->front0.p;
    a.front1 = ctx->front1.p;
    a.inc = ctx->inc.p;
    a.eoff = ctx->eoff.p;
    a.block_sum = ctx->block_sum.p;
    a.ctl = ctx->ctl.p;
    a.rmax = m->rmax;
    a.source = m->source;
    a.edges = m->edges;
    a.vertices = m->vertices;
    a.levels = m->levels;
    a.lastlvl = m->lastlvl;
    a.front_cap = (u32)ctx->front0.cap;
    a.max_levels = 1u << 20;
    a.level_base = ctx->level_base;
    a.trace = ctx->trace_on ? ctx->trace.p : nullptr;
    a.trace_cap = ctx->trace_on ? 4096 : 0;
    a.tile_max = getenv("FORA_TILE_MAX") ? (u32)atoi(getenv("FORA_TILE_MAX")) : TILE_MAX;
    a.l2_hints = getenv("FORA_L2_HINTS") ? (u32)atoi(getenv("FORA_L2_HINTS")) : 1u;
    a.log_v = ctx->log_cap ? ctx->log_v.p : nullptr;
    a.log_r = ctx->log_r.p;
    a.log_cap = (u32)ctx->log_cap;
    a.log_cur = ctx->log_cur.p;
    a.colx = ctx->g.deg_shift ? ctx->g.out_colx : nullptr;
    a.deg_shift = ctx->g.deg_shift;
    a.err = &m->push_err;
    // dense slot-levels (RED + scan): frontier size, as a fraction of the vertices, from which a slot-level takes that path; < 0: never
    {
        const double f = getenv("FORA_PUSH_DENSE") ? atof(getenv("FORA_PUSH_DENSE")) : -1.0;
        a.dense_min = f < 0 ? 0xffffffffu : (u32)std::max(1.0, f * (double)ctx->g.n);
    }
    // edge lists of dense slot-levels (FORA_PUSH_EL=0: dense slot-levels go through the tiles with RED)
    a.el = nullptr; a.el_cap = 0; a.el_count = nullptr; a.el_bad = nullptr;
    a.el_prefetch = getenv("FORA_EL_PF") ? (u32)atol(getenv("FORA_EL_PF")) : (1u << 20);
    a.debug_el_skip = getenv("FORA_DEBUG_EL_SKIP") ? (u32)atol(getenv("FORA_DEBUG_EL_SKIP")) : 0u;
    a.debug_el = getenv("FORA_DEBUG_EL") ? (u32)atoi(getenv("FORA_DEBUG_EL")) : 0u;
    if (a.dense_min != 0xffffffffu && ctx->g.off32 && ctx->el_cap) {
        a.el = ctx->el.p; a.el_cap = (u32)ctx->el_cap; a.el_count = ctx->el_ctl.p; a.el_bad = ctx->el_ctl.p + 2 * MAX_SLOTS;
    }
    // lockstep phase B (FORA_PUSH_LOCKSTEP = group size in units of n edges, e.g. 0.125; unset / 0: off)
    {
        const double f = getenv("FORA_PUSH_LOCKSTEP") ? atof(getenv("FORA_PUSH_LOCKSTEP")) : 0.0;
        a.lockstep_edges = f > 0 ? (u64)std::max(1.0, f * (double)ctx->g.n) : 0;
    }
    return a;
}

// add the logged reserve credits (alpha * r per push) to the reserve vectors and empty the log
static int apply_push_log(fora_ctx* ctx) {
    if (!ctx->log_cap) return FORA_OK;
    apply_log_kernel<<<dim3(ctx->num_sms * 8, ctx->slots), 256, 0, ctx->stream>>>(ctx->g.n, ctx->p.alpha, ctx->log_cur.p, (u32)ctx->log_cap, ctx->log_v.p,
                                                                               ctx->log_r.p, ctx->reserve.p);
    CKL();
    CK(cudaMemsetAsync(ctx->log_cur.p, 0, sizeof(u32) * MAX_SLOTS, ctx->stream));
    return FORA_OK;
}

// launch the persistent push kernel over whatever frontier is in front0 / ctl->fcount[0].  defer_log: leave the reserve
// credits of this launch in the log (the caller applies them once after the last round of the wave).
static int launch_push2(fora_ctx* ctx, bool defer_log);
static int launch_push3(fora_ctx* ctx, bool defer_log);
static int launch_push(fora_ctx* ctx, bool defer_log = false) {
    if (ctx->push_v == 2) return launch_push2(ctx, defer_log);
    if (ctx->push_v == 3) return launch_push3(ctx, defer_log);
    PushArgs a = make_push_args(ctx);
    ctx->level_base += (1u << 20);
    int wrc = set_l2_window(ctx, ctx->win_push_off, ctx->win_push_bytes);
    if (wrc) return wrc;
    if (ctx->push_carve && (wrc = set_persist_limit(ctx, ctx->push_carve))) return wrc;
    kev_begin(ctx, 0);
    if (ctx->g.off32) {
        CsrView<u32> v{ctx->hot_ptr32, ctx->g.out_col};
        void* args[] = {&a, &v};
        CK(cudaLaunchCooperativeKernel(a.dense_min != 0xffffffffu ? (void*)push_kernel<u32, true> : (void*)push_kernel<u32, false>, dim3(ctx->push_grid),
                                       dim3(PUSH_THREADS), args, sizeof(PushSmem<u32>), ctx->stream));
    } else {
        CsrView<int64_t> v{ctx->g.out_ptr64, ctx->g.out_col};
        void* args[] = {&a, &v};
        CK(cudaLaunchCooperativeKernel(a.dense_min != 0xffffffffu ? (void*)push_kernel<int64_t, true> : (void*)push_kernel<int64_t, false>, dim3(ctx->push_grid),
                                       dim3(PUSH_THREADS), args, sizeof(PushSmem<int64_t>), ctx->stream));
    }
    kev_end(ctx);
    ctx->launches++;
    return defer_log ? FORA_OK : apply_push_log(ctx);
}

// Third-generation push (push3.cuh): the same persistent cooperative launch as launch_push, the grid sweeps the slots in lockstep.
static int launch_push3(fora_ctx* ctx, bool defer_log) {
    PushArgs a = make_push_args(ctx);
    P3Args x{};
    x.c = ctx->p3ctl.p;
    x.hub_list = ctx->p3hub.p;
    x.hub_cap = (u32)ctx->p3hub_cap;
    x.est_deg = (u32)std::max<int64_t>(1, (ctx->g.n_edges + ctx->g.n - 1) / std::max(1, ctx->g.n));
    x.budget_sectors = (u32)std::min<double>(4.0e9, std::max(1.0, ctx->p3_budget * (double)ctx->g.n / 4.0));
    x.beg32 = ctx->eoff.p;
    x.hub_deg = ctx->p3_hub_deg;
    x.debug_mode = getenv("FORA_DEBUG_P3_MODE") ? (u32)atoi(getenv("FORA_DEBUG_P3_MODE")) : 0u;
    x.debug_skip_hot = getenv("FORA_DEBUG_P3_SKIPHOT") ? (u32)atoi(getenv("FORA_DEBUG_P3_SKIPHOT")) : 0u;
    x.dense_min = ctx->p3_dense < 0 ? 0xffffffffu : (u32)std::max(1.0, ctx->p3_dense * (double)ctx->g.n);
    x.hub_piece = ctx->p3_hub_piece;
    ctx->level_base += (1u << 20);
    int wrc = set_l2_window(ctx, ctx->win_push_off, ctx->win_push_bytes);
    if (wrc) return wrc;
    kev_begin(ctx, 0);
    if (ctx->g.off32) {
        CsrView<u32> v{ctx->hot_ptr32, ctx->g.out_col};
        void* args[] = {&a, &v, &x};
        CK(cudaLaunchCooperativeKernel((void*)push3_kernel<u32>, dim3(ctx->push3_grid), dim3(P3_THREADS), args, sizeof(P3Smem), ctx->stream));
    } else {
        CsrView<int64_t> v{ctx->g.out_ptr64, ctx->g.out_col};
        void* args[] = {&a, &v, &x};
        CK(cudaLaunchCooperativeKernel((void*)push3_kernel<int64_t>, dim3(ctx->push3_grid), dim3(P3_THREADS), args, sizeof(P3Smem), ctx->stream));
    }
    kev_end(ctx);
    ctx->launches++;
    return defer_log ? FORA_OK : apply_push_log(ctx);
}

// Second-generation push of one round over whatever frontier is installed (push2.cuh): the tail kernel runs the small levels of
// every slot concurrently (one CTA per slot), the sub-wave kernel runs the large levels of push_sub slots at a time with their
// residue vectors L2-resident; repeat until every frontier is empty (normally: head -> sub-waves -> tail, one pass).
static int launch_push2(fora_ctx* ctx, bool defer_log) {
    Push2Args a{};
    a.p = make_push_args(ctx);
    SlotMeta* m = ctx->meta.p;
    a.tail_nf = ctx->tail_nf;
    a.tail_e = ctx->tail_e;
    a.rv = ctx->inc.p;
    a.begs = ctx->front_begs.p;
    a.hub = ctx->hubbuf.p;
    a.hub_cnt = ctx->hubbuf.p + 2 * P2_HUB_CAP;
    a.force = m->force;
    a.left = m->left;
    a.err = &m->push_err;
    a.prefetch = ctx->push_prefetch;
    const int S = ctx->slots, K = ctx->push_sub;
    int wrc = set_l2_window(ctx, ctx->win_push_off, ctx->win_push_bytes);
    if (wrc) return wrc;
    if (ctx->push_carve && (wrc = set_persist_limit(ctx, ctx->push_carve))) return wrc;
    auto tail = [&]() -> int {
        a.slot0 = 0;
        a.k = S;
        if (ctx->g.off32) {
            CsrView<u32> v{ctx->hot_ptr32, ctx->g.out_col};
            push_tail_kernel<u32><<<S, TAIL_THREADS, sizeof(TailSmem), ctx->stream>>>(a, v);
        } else {
            CsrView<int64_t> v{ctx->g.out_ptr64, ctx->g.out_col};
            push_tail_kernel<int64_t><<<S, TAIL_THREADS, sizeof(TailSmem), ctx->stream>>>(a, v);
        }
        CKL();
        return FORA_OK;
    };
    int rc;
    kev_begin(ctx, 0);
    if ((rc = tail())) return rc;
    kev_end(ctx);
    for (int rep = 0;; ++rep) {
        if ((rc = meta_d2h_sync(ctx))) return rc; // which slots still hold a frontier (mapped memory, no copy engine)
        if (ctx->h_meta->push_err) return ctx->fail(FORA_ECUDA, "push: level cap reached");
        bool any = false;
        for (int s = 0; s < S; ++s) any = any || ctx->h_meta->left[s] != 0;
        if (!any) break;
        if (rep >= 4096) return ctx->fail(FORA_ECUDA, "push: the frontier does not drain");
        kev_begin(ctx, 2);
        for (int s0 = 0; s0 < S; s0 += K) {
            const int k = std::min(K, S - s0);
            bool need = false;
            for (int s = s0; s < s0 + k; ++s) need = need || ctx->h_meta->left[s] != 0;
            if (!need) continue;
            a.slot0 = s0;
            a.k = k;
            if (ctx->push_carve && ctx->push_win && ctx->push_v == 2) { // this sub-wave's residue vectors become the persisting part of the L2
                int w2 = set_l2_window(ctx, (size_t)((unsigned char*)(ctx->residue.p + (size_t)s0 * ctx->g.n) - ctx->arena.p), sizeof(double) * (size_t)k * ctx->g.n, ctx->push_carve);
                if (w2) return w2;
            }
            if (ctx->g.off32) {
                CsrView<u32> v{ctx->hot_ptr32, ctx->g.out_col};
                void* args[] = {&a, &v};
                CK(cudaLaunchCooperativeKernel((void*)push2_kernel<u32>, dim3(ctx->push2_grid), dim3(P2_THREADS), args, sizeof(P2Smem), ctx->stream));
            } else {
                CsrView<int64_t> v{ctx->g.out_ptr64, ctx->g.out_col};
                void* args[] = {&a, &v};
                CK(cudaLaunchCooperativeKernel((void*)push2_kernel<int64_t>, dim3(ctx->push2_grid), dim3(P2_THREADS), args, sizeof(P2Smem), ctx->stream));
            }
            ctx->launches++;
        }
        if ((rc = tail())) return rc;
        kev_end(ctx);
    }
    if (ctx->push_carve && ctx->push_win_reset) cudaCtxResetPersistingL2Cache();
    return defer_log ? FORA_OK : apply_push_log(ctx);
}

// rsum / nnz of every slot; with seed_next also the seed lists of the next round at h_meta->next_rmax
static int launch_residue_stats(fora_ctx* ctx, bool seed_next = false) {
    const int S = ctx->slots;
    SlotMeta* m = ctx->meta.p;
    if (seed_next) CK(cudaMemsetAsync(m->seed_count, 0, sizeof(u32) * MAX_SLOTS, ctx->stream)); // next_rmax was uploaded by push_round_active
    residue_partial_kernel<<<dim3(ctx->red_blocks, S), RED_THREADS, 0, ctx->stream>>>(ctx->g.n, ctx->residue.p, ctx->part_sum.p, ctx->part_nnz.p, ctx->hot_deg,
                                                                                     seed_next ? m->next_rmax : nullptr, ctx->front0.p, m->seed_count);
    CKL();
    residue_final_kernel<<<S, 32, 0, ctx->stream>>>(ctx->red_blocks, ctx->part_sum.p, ctx->part_nnz.p, ctx->meta.p->rsum, ctx->meta.p->nnz);
    CKL();
    return FORA_OK;
}

// zero dense state of the first `cnt` slots and install the sources (h_meta->source / qid filled by caller)
static int init_wave(fora_ctx* ctx, int cnt, int seed_source, const int32_t* d_sources) {
    const size_t n = (size_t)ctx->g.n;
    SlotMeta* h = ctx->h_meta;
    for (int s = 0; s < MAX_SLOTS; ++s) {
        if (s >= cnt) h->source[s] = -1;
        h->state[s] = 0; h->active[s] = 0; h->edges[s] = h->vertices[s] = h->levels[s] = 0;
        h->lastlvl[s] = 0; h->rsum[s] = 0; h->nnz[s] = h->nsrc[s] = h->nwalk[s] = h->hops[s] = h->idx_hits[s] = 0;
        h->rmax[s] = ctx->p.rmax;
        h->force[s] = h->left[s] = 0;
    }
    h->push_err = 0;
    ctx->level_base = 0;
    int rc = meta_h2d(ctx);
    if (rc) return rc;
    if (d_sources) {
        if (ctx->g.relabeled) {
            map_ids_kernel<<<1, 64, 0, ctx->stream>>>((size_t)cnt, d_sources, ctx->g.old2new, ctx->meta.p->source);
            CKL();
        } else CK(cudaMemcpyAsync(ctx->meta.p->source, d_sources, sizeof(int32_t) * cnt, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CK(cudaMemsetAsync(ctx->reserve.p, 0, sizeof(double) * n * cnt, ctx->stream));
    CK(cudaMemsetAsync(ctx->residue.p, 0, sizeof(double) * n * cnt, ctx->stream));
    if (ctx->log_cur.p) CK(cudaMemsetAsync(ctx->log_cur.p, 0, sizeof(u32) * MAX_SLOTS, ctx->stream)); // a new wave starts with an empty credit log
    CK(cudaMemsetAsync(ctx->ctl.p, 0, sizeof(PushCtl), ctx->stream));
    push_init_kernel<<<1, MAX_SLOTS, 0, ctx->stream>>>(ctx->g.n, ctx->slots, ctx->meta.p->source, ctx->hot_deg, ctx->reserve.p,
                                                      ctx->residue.p, ctx->front0.p, ctx->ctl.p, seed_source, ctx->meta.p->state);
    CKL();
    return FORA_OK;
}

// one resumable round over the slots flagged in h_meta->active (rmax per slot in h_meta->rmax)
// have_seeds: the previous stats pass already produced this round's seed lists (h_meta->seed_count, front0);
// next_seed: let this round's stats pass prepare the following round at h_meta->next_rmax
static int push_round_active(fora_ctx* ctx, bool have_seeds = false, bool next_seed = false, bool defer_log = false) {
    SlotMeta* h = ctx->h_meta;
    CK(cudaMemcpyAsync(ctx->meta.p->rmax, h->rmax, sizeof(double) * MAX_SLOTS, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->meta.p->active, h->active, sizeof(int32_t) * MAX_SLOTS, cudaMemcpyHostToDevice, ctx->stream));
    // before the push: launch_push2 mirrors the device block back into h_meta while it drains the frontier
    if (next_seed) CK(cudaMemcpyAsync(ctx->meta.p->next_rmax, h->next_rmax, sizeof(double) * MAX_SLOTS, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->ctl.p, 0, sizeof(PushCtl), ctx->stream));
    if (have_seeds) {
        u32 cnt[MAX_SLOTS];
        for (int s = 0; s < MAX_SLOTS; ++s) cnt[s] = h->active[s] ? h->seed_count[s] : 0; // slots that stopped keep their residue
        CK(cudaMemcpyAsync(ctx->ctl.p->fcount[0], cnt, sizeof(cnt), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream)); // cnt lives on this stack frame
    } else {
        const int gx = std::max(1, std::min(ctx->num_sms * 8, (ctx->g.n + 255) / 256));
        push_seed_kernel<<<dim3(gx, ctx->slots), 256, 0, ctx->stream>>>(ctx->g.n, ctx->hot_deg, ctx->residue.p, ctx->meta.p->rmax,
                                                                       ctx->meta.p->active, ctx->front0.p, ctx->ctl.p);
        CKL();
    }
    int rc = launch_push(ctx, defer_log);
    if (rc) return rc;
    return launch_residue_stats(ctx, next_seed);
}

// Push phase of a wave of `cnt` FORA queries: plain (algo.h:954) or --balanced (query.h:848-884).
static int push_wave(fora_ctx* ctx, int cnt, const int32_t* d_sources, double* final_rmax /*[cnt]*/, u64* rounds /*[cnt]*/) {
    const fora_params& p = ctx->p;
    SlotMeta* h = ctx->h_meta;
    int rc;
    if (!p.balanced) {
        if ((rc = init_wave(ctx, cnt, 1, d_sources))) return rc;
        if ((rc = launch_push(ctx))) return rc;
        if ((rc = launch_residue_stats(ctx))) return rc;
        for (int s = 0; s < cnt; ++s) { final_rmax[s] = p.rmax; rounds[s] = 1; }
        return FORA_OK;
    }
    if ((rc = init_wave(ctx, cnt, 0, d_sources))) return rc;
    if ((rc = meta_d2h_sync(ctx))) return rc; // slot states (dangling sources)
    std::vector<double> rmax(cnt, p.rmax * 8), used(cnt, 0.0), rsum(cnt, 1.0);
    std::vector<char> done(cnt, 0);
    std::vector<u64> e0(cnt, 0), v0(cnt, 0), l0(cnt, 0);
    for (int s = 0; s < cnt; ++s) {
        rounds[s] = 0;
        if (h->state[s] != 1) { done[s] = 1; final_rmax[s] = p.rmax; }
    }
    for (int iter = 0; iter < 64; ++iter) {
        bool any = false;
        for (int s = 0; s < cnt; ++s) {
            h->active[s] = 0;
            if (done[s]) continue;
            // estimated_random_walk_cost, query.h:826-839
            double est;
            // shared walks: a walk costs a pool lookup plus its share of the pool build, ~0.2 of a private walk at 48 slots
            // (measured optimum of the loop: FORA_COST_WALK 1.2e-11 .. 3e-11 instead of 6.5e-11, profiles/r2g_experiments.txt)
            const double cw = p.cost_walk * ((ctx->shared_walks && !p.with_idx) ? ctx->shared_cost_scale : 1.0);
            if (!p.with_idx || rmax[s] >= p.rmax) est = p.omega * rsum[s] * (1 - p.alpha) * cw;
            else est = p.omega * rsum[s] * (1 - p.alpha) * (p.cost_walk / 140);
            if (!(est > used[s])) {
                done[s] = 1;
                final_rmax[s] = rmax[s] * 2; // query.h:878
                continue;
            }
            h->active[s] = 1;
            h->rmax[s] = rmax[s];
            e0[s] = h->edges[s]; v0[s] = h->vertices[s]; l0[s] = h->levels[s];
            any = true;
        }
        if (!any) break;
        // the stats pass of this round also lists the seeds of the next one (rmax/2) for every active slot
        for (int s = 0; s < MAX_SLOTS; ++s) h->next_rmax[s] = (s < cnt && h->active[s]) ? rmax[s] / 2 : 0.0;
        if ((rc = push_round_active(ctx, iter > 0, true, true))) return rc; // reserve credits stay logged until the last round
        if ((rc = meta_d2h_sync(ctx))) return rc;
        for (int s = 0; s < cnt; ++s) {
            if (!h->active[s]) continue;
            used[s] += p.cost_edge * (double)(h->edges[s] - e0[s]) + p.cost_vertex * (double)(h->vertices[s] - v0[s]) +
                       p.cost_level * (double)(h->levels[s] - l0[s]);
            rsum[s] = h->rsum[s];
            rmax[s] /= 2;
            rounds[s]++;
        }
    }
    for (int s = 0; s < cnt; ++s)
        if (!done[s]) final_rmax[s] = rmax[s] * 2;
    return apply_push_log(ctx); // one pass per wave: a vertex pushed in several rounds costs one sector round trip, not several
}

// Walk phase of a wave: plan + walk kernels; ppr is accumulated in place into `ppr` ([slots*n],
// holding the reserve on entry).  round_tag distinguishes the Philox streams of top-k rounds.
// groups > 1 launches the walk kernel once per group of slots and calls after_group(lo, hi) behind each launch, so a
// caller can ship finished slots while the remaining ones still walk.
// ---------------------------------------------------------------------------------------------
// Shared walks (opt-in, fora_ctx_set_shared_walks): the wave's queries draw their walks from ONE pool.  Walk j from vertex v is
// computed once per wave and serves every query of the wave that needs at least j + 1 walks from v -- the reference's own
// --with_idx semantics (query.h:290-307: every query reads the same stored destinations), with the index sized by what the wave
// needs (max over its slots of n_v), rebuilt with fresh randomness for every wave and never stored.  Each query's estimate keeps
// its distribution (its walks are independent samples of the walk from v); estimates of queries of the SAME wave are correlated.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sw_max_kernel(int32_t n, const int32_t* __restrict__ srcs, const u64* __restrict__ woff,
                                                     const u64* __restrict__ nsrc, const int32_t* __restrict__ slot_state, u32* __restrict__ cnt,
                                                     u64* __restrict__ overflow) {
    const int slot = blockIdx.y;
    if (slot_state[slot] != 1) return;
    const u64 ns = nsrc[slot];
    const int32_t* __restrict__ sv = srcs + (size_t)slot * n;
    const u64* __restrict__ wo = woff + (size_t)slot * (n + 1);
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < ns; i += (u64)gridDim.x * blockDim.x) {
        const u64 c = wo[i + 1] - wo[i];
        if (c > 0xffffffffull) *overflow = 1; // (more than 2^32 walks from one vertex: the wave keeps its private walks)
        atomicMax(&cnt[sv[i]], (u32)min(c, (u64)0xffffffffu));
    }
}
__global__ void __launch_bounds__(256) sw_flag_kernel(int32_t n, const u32* __restrict__ cnt, u64* __restrict__ cnt64, u32* __restrict__ flag) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
        cnt64[v] = cnt[v];
        flag[v] = cnt[v] != 0;
    }
}
__global__ void __launch_bounds__(256) sw_compact_kernel(int32_t n, const u32* __restrict__ cnt, const u32* __restrict__ pos,
                                                         const u64* __restrict__ off, int32_t* __restrict__ srcs, u64* __restrict__ woff,
                                                         u64* __restrict__ totals /* [2]: sources, walks */) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
        if (cnt[v]) {
            srcs[pos[v]] = v;
            woff[pos[v]] = off[v];
        }
        if (v == n - 1) {
            const u64 ns = pos[v] + (cnt[v] != 0), nw = off[v] + cnt[v];
            woff[ns] = nw;
            totals[0] = ns;
            totals[1] = nw;
        }
    }
}
struct BulkPlan;
static int launch_bulk(fora_ctx* ctx, const BulkPlan& bp, u64* hops_out, int parts, const std::function<int(int, u64, u64)>& after_part);
static int build_shared_walks(fora_ctx* ctx, int no_zero_hop, bool* usable);

static int walk_wave(fora_ctx* ctx, double* ppr, int per_round, int opt, int no_zero_hop, u32 round_tag, const u64* idx_used,
                     u32 part = 0, u32 nparts = 1, int groups = 1, const std::function<int(int, int)>& after_group = nullptr) {
    const DeviceGraph& g = ctx->g;
    const int S = ctx->slots;
    SlotMeta* m = ctx->meta.p;
    PlanArgs pa{};
    pa.n = g.n; pa.alpha = ctx->p.alpha; pa.omega = ctx->p.omega; pa.opt = opt; pa.per_round = per_round;
    pa.residue = ctx->residue.p; pa.ppr = ppr; pa.rsum = m->rsum; pa.slot_state = m->state;
    pa.blk_src = ctx->blk_src.p; pa.blk_walk = ctx->blk_walk.p; pa.srcs = ctx->srcs.p; pa.woff = ctx->woff.p;
    pa.incs = ctx->incs.p; pa.nsrc = m->nsrc; pa.nwalk = m->nwalk;
    pa.nblk = (g.n + PLAN_THREADS * PLAN_VPT - 1) / (PLAN_THREADS * PLAN_VPT);
    pa.no_credit = part > 0;
    plan_kernel<false><<<dim3(pa.nblk, S), PLAN_THREADS, 0, ctx->stream>>>(pa);
    CKL();
    plan_scan_kernel<<<S, 1024, 0, ctx->stream>>>(pa);
    CKL();
    plan_kernel<true><<<dim3(pa.nblk, S), PLAN_THREADS, 0, ctx->stream>>>(pa);
    CKL();
    const int cgx = std::max<int>(1, std::min<size_t>((size_t)ctx->num_sms * 4, (ctx->chunk_cap + 255) / 256));
    chunk_start_kernel<<<dim3(cgx, S), 256, 0, ctx->stream>>>(g.n, ctx->woff.p, m->nsrc, m->nwalk, ctx->chunk_first.p, ctx->chunk_cap, m->state);
    CKL();
    WalkArgs wa{};
    wa.n = g.n;
    wa.alpha_thr = (u32)std::min(4294967295.0, ctx->p.alpha * 4294967296.0);
    wa.seed_lo = (u32)ctx->seed; wa.seed_hi = (u32)(ctx->seed >> 32);
    wa.with_idx = ctx->p.with_idx && ctx->has_index;
    const bool shared = ctx->shared_walks && !wa.with_idx && per_round == 0 && nparts == 1 && !idx_used;
    bool shared_ok = false;
    if (shared) {
        int brc = build_shared_walks(ctx, no_zero_hop, &shared_ok);
        if (brc) return brc;
    }
    wa.srcs = ctx->srcs.p; wa.woff = ctx->woff.p; wa.incs = ctx->incs.p; wa.nsrc = m->nsrc; wa.nwalk = m->nwalk;
    wa.chunk_first = ctx->chunk_first.p; wa.chunk_cap = ctx->chunk_cap; wa.slot_state = m->state; wa.qid = m->qid;
    wa.part = part; wa.nparts = nparts;
    wa.round_tag = round_tag; wa.ppr = ppr; wa.hops = m->hops; wa.idx_hits = m->idx_hits;
    wa.idx_off = ctx->idx_off.p; wa.idx_cnt = ctx->idx_cnt.p; wa.idx_dest = ctx->idx_dest.p; wa.idx_used = idx_used;
    if (shared && shared_ok) { // every walk of the wave is a hit in the pool just built
        wa.with_idx = 1;
        wa.all_idx = 1;
        wa.idx_off = ctx->sw_off.p; wa.idx_cnt = ctx->sw_cnt64.p; wa.idx_dest = ctx->sw_dest.p; wa.idx_used = nullptr;
    }
    // neighbour slots beyond the first 32 MB of the (hot-first) column array stream through the L2 with evict_first
    // (measured: 16..64 MB within 1 %, +2.6 % hops/s over no hint)
    wa.hot_elems = (u64)((getenv("FORA_WALK_HOT_MB") ? atof(getenv("FORA_WALK_HOT_MB")) : 32.0) * 262144.0);
    wa.debug_no_red = getenv("FORA_DEBUG_NO_RED") ? atoi(getenv("FORA_DEBUG_NO_RED")) : 0;
    const int wgx = ctx->num_sms * (getenv("FORA_WALK_GRID") ? atoi(getenv("FORA_WALK_GRID")) : 16);
    {
        int lrc = set_persist_limit(ctx, ctx->persist_walk);
        if (lrc) return lrc;
    }
    if (ppr == ctx->reserve.p) {
        int wrc = set_l2_window(ctx, ctx->win_walk_off, ctx->win_walk_bytes);
        if (wrc) return wrc;
    }
    if (getenv("FORA_WALK_GROUPS")) groups = std::max(groups, atoi(getenv("FORA_WALK_GROUPS"))); // development knob: slots per launch = S / groups
    groups = std::max(1, std::min(groups, S));
    for (int gi = 0; gi < groups; ++gi) {
        const int lo = (int)((long long)S * gi / groups), hi = (int)((long long)S * (gi + 1) / groups);
        if (hi <= lo) continue;
        wa.slot0 = lo;
        const dim3 grid(wgx, hi - lo);
        kev_begin(ctx, 1);
        if (wa.all_idx) { // shared walks: lookups in the wave's pool, nothing is walked
            if (g.off32) {
                CsrView<u32> v{ctx->hot_ptr32, g.out_col};
                walk_kernel<u32, false, false, OUT_POOL><<<grid, WALK_THREADS, 0, ctx->stream>>>(wa, v);
            } else {
                CsrView<int64_t> v{g.out_ptr64, g.out_col};
                walk_kernel<int64_t, false, false, OUT_POOL><<<grid, WALK_THREADS, 0, ctx->stream>>>(wa, v);
            }
        } else if (g.off32) {
            CsrView<u32> v{ctx->hot_ptr32, g.out_col};
            if (wa.hot_elems && wa.hot_elems < (u64)g.n_edges) {
                if (no_zero_hop) walk_kernel<u32, true, true><<<grid, WALK_THREADS, 0, ctx->stream>>>(wa, v);
                else walk_kernel<u32, false, true><<<grid, WALK_THREADS, 0, ctx->stream>>>(wa, v);
            } else {
                if (no_zero_hop) walk_kernel<u32, true, false><<<grid, WALK_THREADS, 0, ctx->stream>>>(wa, v);
                else walk_kernel<u32, false, false><<<grid, WALK_THREADS, 0, ctx->stream>>>(wa, v);
            }
        } else {
            CsrView<int64_t> v{g.out_ptr64, g.out_col};
            if (no_zero_hop) walk_kernel<int64_t, true, false><<<grid, WALK_THREADS, 0, ctx->stream>>>(wa, v);
            else walk_kernel<int64_t, false, false><<<grid, WALK_THREADS, 0, ctx->stream>>>(wa, v);
        }
        kev_end(ctx);
        CKL();
        if (after_group) {
            int arc = after_group(lo, hi);
            if (arc) return arc;
        }
    }
    return FORA_OK;
}

// Top-k select of `nsl` dense vectors base + stride*slots[j] in four launches (topk.cuh).  Afterwards the (optionally sorted)
// output list of vector j is ctx->sel_on.p / sel_ov.p + j*ctx->sel_p2, h_res[j] holds its k-th key and entry count.
static int select_batch(fora_ctx* ctx, const double* base, size_t stride, const int* slots, int nsl, u32 k, bool sort, SelResult* h_res) {
    if (nsl <= 0) return FORA_OK;
    if (nsl > MAX_SLOTS) return ctx->fail(FORA_EINVAL, "select_batch: too many vectors");
    const int32_t n = ctx->g.n;
    u32 p2 = 1;
    while (p2 < k) p2 <<= 1;
    CK(ctx->sel_st.ensure(MAX_SLOTS));
    CK(ctx->sel_res.ensure(MAX_SLOTS));
    CK(ctx->sel_slots.ensure(MAX_SLOTS));
    CK(ctx->sel_on.ensure((size_t)p2 * MAX_SLOTS));
    CK(ctx->sel_ov.ensure((size_t)p2 * MAX_SLOTS));
    ctx->sel_p2 = p2;
    const size_t cand = (size_t)n * (size_t)std::max(nsl, ctx->slots); // worst case: every entry shares the k-th entry's exponent
    CK(ctx->sel_ck.ensure(cand));
    CK(ctx->sel_ci.ensure(cand));
    int32_t ids[MAX_SLOTS];
    for (int j = 0; j < nsl; ++j) ids[j] = slots[j];
    CK(cudaMemcpyAsync(ctx->sel_slots.p, ids, sizeof(int32_t) * nsl, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->sel_st.p, 0, sizeof(SelSlot) * nsl, ctx->stream));
    const int gx = std::max(1, std::min(std::max(ctx->num_sms * 16 / nsl, 8), (n + 255) / 256));
    sel_hist_kernel<<<dim3(gx, nsl), 256, 0, ctx->stream>>>(base, stride, n, ctx->sel_slots.p, ctx->sel_st.p);
    CKL();
    sel_pick_kernel<<<nsl, 256, 0, ctx->stream>>>(ctx->sel_st.p, k);
    CKL();
    sel_classify_kernel<<<dim3(gx, nsl), 256, 0, ctx->stream>>>(base, stride, n, ctx->sel_slots.p, ctx->sel_st.p, ctx->sel_ck.p, ctx->sel_ci.p, (size_t)n,
                                                               ctx->sel_on.p, ctx->sel_ov.p, p2);
    CKL();
    sel_finish_kernel<<<nsl, 1024, 0, ctx->stream>>>(ctx->sel_st.p, ctx->sel_res.p, ctx->sel_ck.p, ctx->sel_ci.p, (size_t)n, ctx->sel_on.p, ctx->sel_ov.p, p2, k, sort ? 1 : 0);
    CKL();
    CK(cudaMemcpyAsync(h_res, ctx->sel_res.p, sizeof(SelResult) * nsl, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FORA_OK;
}
static inline double sel_kth(const SelResult& r) { // kth_ppr(): 0 when there are fewer than k positive entries
    double t = 0.0;
    if (!r.all_positive) memcpy(&t, &r.key_T, sizeof t);
    return t;
}

static void fill_stat(fora_ctx* ctx, int s, double final_rmax, u64 rounds, fora_query_stat* st) {
    const SlotMeta* h = ctx->h_meta;
    memset(st, 0, sizeof *st);
    st->rsum = (h->state[s] == 1 || h->state[s] == 3) ? h->rsum[s] : 0.0;
    st->final_rmax = final_rmax;
    st->n_walks = h->nwalk[s];
    st->n_idx_hits = h->idx_hits[s];
    st->walk_hops = h->hops[s];
    st->edges_pushed = h->edges[s];
    st->vertices_pushed = h->vertices[s];
    st->push_levels = h->levels[s];
    st->push_rounds = rounds;
    st->n_sources = h->nsrc[s];
}

// =============================================================================================
// push test hooks
// =============================================================================================
static int require_ready(fora_ctx* ctx, double omega_max) {
    if (!ctx) return FORA_EINVAL;
    if (!ctx->params_set) return ctx->fail(FORA_EINVAL, "fora_params_set has not been called");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return ctx->fail(FORA_ECUDA, cudaGetErrorString(e));
    return ensure_slots(ctx, omega_max);
}

static int download_fwd(fora_ctx* ctx, int slot, double* reserve, double* residue) {
    const size_t n = (size_t)ctx->g.n;
    int rc;
    CK(ctx->scratchd.ensure(n));
    if (reserve) {
        if ((rc = vec_to_original(ctx, ctx->reserve.p + n * slot, ctx->scratchd.p, n))) return rc;
        CK(cudaMemcpyAsync(reserve, ctx->scratchd.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (residue) {
        if ((rc = vec_to_original(ctx, ctx->residue.p + n * slot, ctx->scratchd.p, n))) return rc;
        CK(cudaMemcpyAsync(residue, ctx->scratchd.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return FORA_OK;
}

extern "C" int fora_push_only(fora_ctx* ctx, int32_t source, double rmax, double* reserve, double* residue, double* rsum,
                              fora_query_stat* stat) {
    int rc = require_ready(ctx, ctx ? ctx->p.omega : 0);
    if (rc) return rc;
    if (source < 0 || source >= ctx->g.n) return ctx->fail(FORA_EINVAL, "source out of range");
    const double keep = ctx->p.rmax;
    const int keep_bal = ctx->p.balanced;
    ctx->p.rmax = rmax;
    ctx->p.balanced = 0;
    ctx->h_meta->source[0] = to_internal(ctx, source);
    ctx->h_meta->qid[0] = 0;
    double fr;
    u64 rounds;
    rc = push_wave(ctx, 1, nullptr, &fr, &rounds);
    ctx->p.rmax = keep;
    ctx->p.balanced = keep_bal;
    if (rc) return rc;
    if ((rc = meta_d2h_sync(ctx))) return rc;
    if (rsum) *rsum = ctx->h_meta->state[0] == 1 ? ctx->h_meta->rsum[0] : 0.0;
    if (stat) fill_stat(ctx, 0, rmax, 1, stat);
    ctx->session_source = -1;
    return download_fwd(ctx, 0, reserve, residue);
}

extern "C" int fora_push_begin(fora_ctx* ctx, int32_t source) {
    int rc = require_ready(ctx, ctx ? ctx->p.omega : 0);
    if (rc) return rc;
    if (source < 0 || source >= ctx->g.n) return ctx->fail(FORA_EINVAL, "source out of range");
    ctx->h_meta->source[0] = to_internal(ctx, source);
    ctx->h_meta->qid[0] = 0;
    if ((rc = init_wave(ctx, 1, 0, nullptr))) return rc;
    if ((rc = meta_d2h_sync(ctx))) return rc;
    ctx->session_source = source;
    return FORA_OK;
}

extern "C" int fora_push_round(fora_ctx* ctx, double rmax, double* reserve, double* residue, double* rsum, fora_query_stat* stat) {
    if (!ctx) return FORA_EINVAL;
    if (ctx->session_source < 0) return ctx->fail(FORA_EINVAL, "fora_push_begin has not been called");
    CK(cudaSetDevice(ctx->device));
    SlotMeta* h = ctx->h_meta;
    int rc;
    if (h->state[0] == 1) {
        for (int s = 0; s < MAX_SLOTS; ++s) h->active[s] = 0;
        h->active[0] = 1;
        h->rmax[0] = rmax;
        if ((rc = push_round_active(ctx))) return rc;
    } else if ((rc = launch_residue_stats(ctx))) return rc;
    if ((rc = meta_d2h_sync(ctx))) return rc;
    if (rsum) *rsum = h->state[0] == 1 ? h->rsum[0] : 0.0;
    if (stat) fill_stat(ctx, 0, rmax, 1, stat);
    return download_fwd(ctx, 0, reserve, residue);
}

// =============================================================================================
// walks test hook / Monte-Carlo building block
// =============================================================================================
// Bulk walks -- index build (build.h:344-354), montecarlo_query / bippr_query (query.h:25-31, 81-88), the random_walk test
// hook -- run through the SAME chunked, warp-converged walk kernel as the query path (walk.cuh), with the destination stored
// at the walk's global index (OUT_DEST) or counted (OUT_COUNT) instead of added to a PPR vector.  The plan is the query
// path's: sources (internal ids) with an exclusive prefix of their walk counts.
struct BulkPlan {
    const int32_t* d_srcs = nullptr; // [nsrc] internal vertex ids, every one with at least one walk
    const u64* d_woff = nullptr;     // [nsrc + 1]
    u64 nsrc = 0, nwalk = 0;
    int32_t* out_dest = nullptr;     // device [nwalk], original ids (OUT_DEST) ...
    u64* out_counts = nullptr;       // ... or device [n] histogram in internal ids (OUT_COUNT)
    u32 key_tag = 0;                 // distinguishes index build / MC / BiPPR / test streams
    int no_zero_hop = 0;
    int internal_ids = 0;            // OUT_DEST: keep the engine's internal vertex ids (a per-wave virtual index)
};
struct BulkMeta { // what walk_kernel reads per slot; slot 0 of a private block
    u64 nsrc, nwalk, hops, idx_hits;
    int32_t state;
    u32 qid;
};
// parts > 1: one launch per contiguous chunk range, after_part(part, first walk, end walk) behind each (e.g. to ship that slice)
static int launch_bulk(fora_ctx* ctx, const BulkPlan& bp, u64* hops_out, int parts = 1,
                       const std::function<int(int, u64, u64)>& after_part = nullptr);
static int launch_bulk(fora_ctx* ctx, const BulkPlan& bp, u64* hops_out, int parts, const std::function<int(int, u64, u64)>& after_part) {
    const DeviceGraph& g = ctx->g;
    if (hops_out) *hops_out = 0;
    if (bp.nwalk == 0) return FORA_OK;
    const u64 CH = walk_chunk_size(bp.nwalk);
    const u64 nchunks = (bp.nwalk + CH - 1) / CH;
    CK(ctx->bulk_chunk_first.ensure(nchunks + 2));
    CK(ctx->bulk_meta.ensure(sizeof(BulkMeta)));
    BulkMeta hm{};
    hm.nsrc = bp.nsrc; hm.nwalk = bp.nwalk; hm.state = 1; hm.qid = bp.key_tag;
    CK(cudaMemcpyAsync(ctx->bulk_meta.p, &hm, sizeof hm, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream)); // hm lives on this stack frame
    BulkMeta* dm = (BulkMeta*)ctx->bulk_meta.p;
    const int cgx = (int)std::max<u64>(1, std::min<u64>((u64)ctx->num_sms * 4, (nchunks + 256) / 256));
    chunk_start_kernel<<<dim3(cgx, 1), 256, 0, ctx->stream>>>(g.n, bp.d_woff, &dm->nsrc, &dm->nwalk, ctx->bulk_chunk_first.p, (size_t)(nchunks + 2), &dm->state);
    CKL();
    WalkArgs wa{};
    wa.n = g.n;
    wa.alpha_thr = (u32)std::min(4294967295.0, ctx->p.alpha * 4294967296.0);
    wa.seed_lo = (u32)ctx->seed; wa.seed_hi = (u32)(ctx->seed >> 32);
    wa.srcs = bp.d_srcs; wa.woff = bp.d_woff; wa.incs = nullptr; wa.nsrc = &dm->nsrc; wa.nwalk = &dm->nwalk;
    wa.chunk_first = ctx->bulk_chunk_first.p; wa.chunk_cap = (size_t)(nchunks + 2); wa.slot_state = &dm->state; wa.qid = &dm->qid;
    wa.round_tag = 0x5bd1e995u; wa.hops = &dm->hops; wa.idx_hits = &dm->idx_hits;
    wa.out_dest = bp.out_dest; wa.out_counts = bp.out_counts; wa.new2old = (g.relabeled && !bp.internal_ids) ? g.new2old : nullptr;
    parts = (int)std::max<u64>(1, std::min<u64>((u64)parts, nchunks));
    wa.nparts = (u32)parts;
    const bool to_dest = bp.out_dest != nullptr;
    for (int part = 0; part < parts; ++part) {
        wa.part = (u32)part;
        const u64 c_lo = parts > 1 ? nchunks * (u64)part / (u64)parts : 0, c_hi = parts > 1 ? nchunks * (u64)(part + 1) / (u64)parts : nchunks;
        const dim3 grid((unsigned)std::max<u64>(1, std::min<u64>((u64)ctx->num_sms * 16, c_hi - c_lo)), 1);
        kev_begin(ctx, 1);
        if (g.off32) {
            CsrView<u32> v{g.out_ptr32, g.out_col};
            if (to_dest) {
                if (bp.no_zero_hop) walk_kernel<u32, true, false, OUT_DEST><<<grid, WALK_THREADS, 0, ctx->stream>>>(wa, v);
                else walk_kernel<u32, false, false, OUT_DEST><<<grid, WALK_THREADS, 0, ctx->stream>>>(wa, v);
            } else {
                if (bp.no_zero_hop) walk_kernel<u32, true, false, OUT_COUNT><<<grid, WALK_THREADS, 0, ctx->stream>>>(wa, v);
                else walk_kernel<u32, false, false, OUT_COUNT><<<grid, WALK_THREADS, 0, ctx->stream>>>(wa, v);
            }
        } else {
            CsrView<int64_t> v{g.out_ptr64, g.out_col};
            if (to_dest) {
                if (bp.no_zero_hop) walk_kernel<int64_t, true, false, OUT_DEST><<<grid, WALK_THREADS, 0, ctx->stream>>>(wa, v);
                else walk_kernel<int64_t, false, false, OUT_DEST><<<grid, WALK_THREADS, 0, ctx->stream>>>(wa, v);
            } else {
                if (bp.no_zero_hop) walk_kernel<int64_t, true, false, OUT_COUNT><<<grid, WALK_THREADS, 0, ctx->stream>>>(wa, v);
                else walk_kernel<int64_t, false, false, OUT_COUNT><<<grid, WALK_THREADS, 0, ctx->stream>>>(wa, v);
            }
        }
        kev_end(ctx);
        CKL();
        if (after_part) {
            int arc = after_part(part, c_lo * CH, std::min(bp.nwalk, c_hi * CH));
            if (arc) return arc;
        }
    }
    if (hops_out) {
        CK(cudaMemcpyAsync(hops_out, &dm->hops, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        kev_harvest(ctx);
    }
    return FORA_OK;
}
// `count` walks from ONE source (internal id): the plan is a single segment
static int launch_bulk_single(fora_ctx* ctx, int32_t source_internal, u64 count, int32_t* out_dest, u64* out_counts, u32 key_tag, int no_zero_hop,
                              u64* hops_out) {
    CK(ctx->bulk_small.ensure(4));
    const u64 hw[4] = {0, count, (u64)(u32)source_internal, 0};
    CK(cudaMemcpyAsync(ctx->bulk_small.p, hw, sizeof hw, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    BulkPlan bp;
    bp.d_srcs = (const int32_t*)(ctx->bulk_small.p + 2); // low word of the third entry (little endian)
    bp.d_woff = ctx->bulk_small.p;
    bp.nsrc = 1; bp.nwalk = count; bp.out_dest = out_dest; bp.out_counts = out_counts; bp.key_tag = key_tag; bp.no_zero_hop = no_zero_hop;
    return launch_bulk(ctx, bp, hops_out);
}

// the wave's walk pool: counts (max over slots), offsets, compacted plan, destinations (see sw_max_kernel)
static int build_shared_walks(fora_ctx* ctx, int no_zero_hop, bool* usable) {
    *usable = false;
    const DeviceGraph& g = ctx->g;
    const size_t n = (size_t)g.n;
    const int S = ctx->slots;
    SlotMeta* m = ctx->meta.p;
    CK(ctx->sw_cnt.ensure(n)); CK(ctx->sw_cnt64.ensure(n)); CK(ctx->sw_off.ensure(n)); CK(ctx->sw_flag.ensure(n)); CK(ctx->sw_pos.ensure(n));
    CK(ctx->sw_srcs.ensure(n)); CK(ctx->sw_woff.ensure(n + 4));
    CK(cudaMemsetAsync(ctx->sw_cnt.p, 0, sizeof(u32) * n, ctx->stream));
    u64* d_tot = ctx->sw_woff.p + n + 1; // three spare words behind the prefix: sources, walks, overflow flag
    CK(cudaMemsetAsync(d_tot, 0, 3 * sizeof(u64), ctx->stream));
    sw_max_kernel<<<dim3(ctx->num_sms * 4, S), 256, 0, ctx->stream>>>(g.n, ctx->srcs.p, ctx->woff.p, m->nsrc, m->state, ctx->sw_cnt.p, d_tot + 2);
    CKL();
    sw_flag_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(g.n, ctx->sw_cnt.p, ctx->sw_cnt64.p, ctx->sw_flag.p);
    CKL();
    size_t t1 = 0, t2 = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, t1, ctx->sw_cnt64.p, ctx->sw_off.p, (int)n, ctx->stream));
    CK(cub::DeviceScan::ExclusiveSum(nullptr, t2, ctx->sw_flag.p, ctx->sw_pos.p, (int)n, ctx->stream));
    size_t tb = std::max(t1, t2);
    CK(ctx->sw_tmp.ensure(tb));
    CK(cub::DeviceScan::ExclusiveSum(ctx->sw_tmp.p, tb, ctx->sw_cnt64.p, ctx->sw_off.p, (int)n, ctx->stream));
    CK(cub::DeviceScan::ExclusiveSum(ctx->sw_tmp.p, tb, ctx->sw_flag.p, ctx->sw_pos.p, (int)n, ctx->stream));
    sw_compact_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(g.n, ctx->sw_cnt.p, ctx->sw_pos.p, ctx->sw_off.p, ctx->sw_srcs.p, ctx->sw_woff.p, d_tot);
    CKL();
    u64 tot[3] = {0, 0, 0};
    CK(cudaMemcpyAsync(tot, d_tot, sizeof tot, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->sw_built_walks = tot[1];
    ctx->sw_built_hops = 0;
    if (tot[2]) return FORA_OK; // a count beyond 32 bits: private walks for this wave
    *usable = true;
    if (tot[1] == 0) return FORA_OK;
    if (ctx->sw_dest.cap < (size_t)tot[1]) CK(ctx->sw_dest.ensure((size_t)tot[1] + (size_t)tot[1] / 2)); // headroom: pools differ from wave to wave, a
                                                                                                        // reallocation is a device synchronisation
    BulkPlan bp;
    bp.d_srcs = ctx->sw_srcs.p; bp.d_woff = ctx->sw_woff.p; bp.nsrc = tot[0]; bp.nwalk = tot[1]; bp.out_dest = ctx->sw_dest.p;
    // one stream per wave: ctx seed x global index of the wave's first query (reproducible; a different cut of the query list into
    // waves gives different pools)
    bp.key_tag = 0x5a000000u ^ ctx->h_meta->qid[0];
    bp.no_zero_hop = no_zero_hop;
    bp.internal_ids = 1;
    ++ctx->sw_waves;
    return launch_bulk(ctx, bp, nullptr, 1, nullptr);
}

extern "C" int fora_random_walks(fora_ctx* ctx, int32_t start, int64_t count, int no_zero_hop, int32_t* dest, uint64_t* hops) {
    if (!ctx || !ctx->g.n) return ctx ? ctx->fail(FORA_EINVAL, "no graph") : FORA_EINVAL;
    if (start < 0 || start >= ctx->g.n || count < 0) return ctx->fail(FORA_EINVAL, "bad start/count");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->scratch32.ensure((size_t)std::max<int64_t>(count, 1)));
    u64 h = 0;
    int rc = launch_bulk_single(ctx, to_internal(ctx, start), (u64)count, ctx->scratch32.p, nullptr, 0x77a1c5u, no_zero_hop, &h);
    if (rc) return rc;
    if (dest && count) CK(cudaMemcpyAsync(dest, ctx->scratch32.p, sizeof(int32_t) * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
    if (hops) *hops = h;
    CK(cudaStreamSynchronize(ctx->stream));
    return FORA_OK;
}

extern "C" int fora_compute_ppr(fora_ctx* ctx, const double* reserve, const double* residue, double rsum, double* ppr,
                                fora_query_stat* stat) {
    int rc = require_ready(ctx, ctx ? ctx->p.omega : 0);
    if (rc) return rc;
    if (!reserve || !residue || !ppr) return ctx->fail(FORA_EINVAL, "null vector");
    const size_t n = (size_t)ctx->g.n;
    { // the walk plan is sized from omega*rsum + n walks: rsum must be the sum of the residues (query.h:255-270 takes it from the push)
        double sum = 0.0;
        for (size_t i = 0; i < n; ++i) {
            if (!(residue[i] >= 0.0)) return ctx->fail(FORA_EINVAL, "compute_ppr: negative or NaN residue");
            sum += residue[i];
        }
        if (!(rsum >= 0.0 && rsum <= 1.0 + 1e-9) || fabs(sum - rsum) > 1e-9 * std::max(1.0, rsum)) return ctx->fail(FORA_EINVAL, "compute_ppr: rsum is not the sum of the residue vector");
    }
    SlotMeta* h = ctx->h_meta;
    memset(h, 0, sizeof *h);
    for (int s = 0; s < MAX_SLOTS; ++s) h->source[s] = -1;
    h->source[0] = 0;
    h->state[0] = rsum == 0.0 ? 2 : 1; // query.h:267-268: rsum == 0 -> ppr = reserve only
    h->rsum[0] = rsum;
    if ((rc = meta_h2d(ctx))) return rc;
    CK(ctx->scratchd.ensure(n));
    CK(cudaMemcpyAsync(ctx->scratchd.p, reserve, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = vec_to_internal(ctx, ctx->scratchd.p, ctx->reserve.p, n))) return rc;
    CK(cudaMemcpyAsync(ctx->scratchd.p, residue, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = vec_to_internal(ctx, ctx->scratchd.p, ctx->residue.p, n))) return rc;
    if ((rc = walk_wave(ctx, ctx->reserve.p, 0, ctx->p.opt, ctx->p.opt, 0, nullptr))) return rc;
    if ((rc = meta_d2h_sync(ctx))) return rc;
    if (stat) fill_stat(ctx, 0, ctx->p.rmax, 0, stat);
    if ((rc = vec_to_original(ctx, ctx->reserve.p, ctx->scratchd.p, n))) return rc;
    CK(cudaMemcpyAsync(ppr, ctx->scratchd.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->session_source = -1;
    return FORA_OK;
}


// =============================================================================================
// backward push / BiPPR
// =============================================================================================
static int ensure_bwd(fora_ctx* ctx) {
    const DeviceGraph& g = ctx->g;
    if (!g.has_in) return ctx->fail(FORA_EINVAL, "backward push needs the in-CSR (fora_graph_upload with in_ptr/in_col)");
    const size_t n = (size_t)g.n;
    if (ctx->bwd_blocks && ctx->bwd_res.cap >= n * ctx->bwd_blocks) return FORA_OK;
    const size_t per_block = 32 * n; // bytes
    int nb = (int)std::min<size_t>((size_t)ctx->num_sms * 4, std::max<size_t>(4, ((size_t)8 << 30) / per_block));
    CK(ctx->bwd_res.ensure(n * nb));
    CK(ctx->bwd_rv.ensure(n * nb));
    CK(ctx->bwd_lists.ensure(4 * n * nb));
    CK(cudaMemsetAsync(ctx->bwd_res.p, 0, sizeof(double) * n * nb, ctx->stream));
    CK(cudaMemsetAsync(ctx->bwd_lists.p, 0, sizeof(int32_t) * 4 * n * nb, ctx->stream)); // stamps start at 0 = never listed
    ctx->bwd_epoch = 1;
    ctx->bwd_blocks = nb;
    return FORA_OK;
}

// run the backward pushes of targets [t_begin, t_end) of one query; counts/ppr may be null (test hook)
static int launch_bwd(fora_ctx* ctx, int32_t source, double rmax, double omega, const u64* counts, double* ppr, int32_t t_begin,
                      int32_t t_end, double* full_reserve, int keep_residue, u64* d_edges, int32_t* d_overflow) {
    const DeviceGraph& g = ctx->g;
    const int nb = keep_residue ? 1 : std::min(ctx->bwd_blocks, std::max(1, t_end - t_begin));
    const u32 span = (u32)std::max(1, t_end - t_begin);
    if (ctx->bwd_epoch > 0xffffffffu - span) { // stamp values would wrap: forget every stamp and start over
        CK(cudaMemsetAsync(ctx->bwd_lists.p, 0, sizeof(int32_t) * 4 * (size_t)g.n * ctx->bwd_blocks, ctx->stream));
        ctx->bwd_epoch = 1;
    }
    const u32 epoch_base = ctx->bwd_epoch;
    ctx->bwd_epoch += span;
    if (g.off32) {
        BwdArgs<u32> a{g.n, ctx->p.alpha, rmax, omega, source, g.in_ptr32, g.in_col, g.deg, counts, ctx->bwd_res.p, ctx->bwd_rv.p,
                       ctx->bwd_lists.p, d_overflow, epoch_base, ppr, t_begin, t_end, full_reserve, keep_residue, d_edges};
        bippr_kernel<u32><<<nb, BWD_THREADS, 0, ctx->stream>>>(a);
    } else {
        BwdArgs<int64_t> a{g.n, ctx->p.alpha, rmax, omega, source, g.in_ptr64, g.in_col, g.deg, counts, ctx->bwd_res.p, ctx->bwd_rv.p,
                           ctx->bwd_lists.p, d_overflow, epoch_base, ppr, t_begin, t_end, full_reserve, keep_residue, d_edges};
        bippr_kernel<int64_t><<<nb, BWD_THREADS, 0, ctx->stream>>>(a);
    }
    CKL();
    return FORA_OK;
}

extern "C" int fora_reverse_push(fora_ctx* ctx, int32_t target, double rmax, double* reserve, double* residue) {
    if (!ctx || !ctx->g.n) return ctx ? ctx->fail(FORA_EINVAL, "no graph") : FORA_EINVAL;
    if (target < 0 || target >= ctx->g.n) return ctx->fail(FORA_EINVAL, "target out of range");
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_bwd(ctx);
    if (rc) return rc;
    const size_t n = (size_t)ctx->g.n;
    CK(ctx->scratchd.ensure(n));
    CK(ctx->scratch32.ensure(4));
    CK(cudaMemsetAsync(ctx->scratchd.p, 0, sizeof(double) * n, ctx->stream));
    CK(cudaMemsetAsync(ctx->scratch32.p, 0, sizeof(int32_t) * 4, ctx->stream));
    const int32_t t_int = to_internal(ctx, target);
    if ((rc = launch_bwd(ctx, -1, rmax, 1.0, nullptr, nullptr, t_int, t_int + 1, ctx->scratchd.p, 1, nullptr, ctx->scratch32.p))) return rc;
    CK(ctx->stage.ensure(n));
    if (reserve) {
        if ((rc = vec_to_original(ctx, ctx->scratchd.p, ctx->stage.p, n))) return rc;
        CK(cudaMemcpyAsync(reserve, ctx->stage.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (residue) {
        if ((rc = vec_to_original(ctx, ctx->bwd_res.p, ctx->stage.p, n))) return rc;
        CK(cudaMemcpyAsync(residue, ctx->stage.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    // the hook left block 0's scratch dirty: wipe it (dense memset, test path only)
    CK(cudaMemsetAsync(ctx->bwd_res.p, 0, sizeof(double) * n, ctx->stream));
    int32_t ovf = 0;
    CK(cudaMemcpyAsync(&ovf, ctx->scratch32.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ovf) { ctx->bwd_blocks = 0; return ctx->fail(FORA_ECUDA, "backward push: touched list overflow"); } // scratch is re-zeroed by the next ensure_bwd
    return FORA_OK;
}

// bippr_query (query.h:71-124) for one source; result in d_ppr (device double[n])
static int bippr_one(fora_ctx* ctx, int32_t source, u32 qid, double* d_ppr, u64* n_walks, u64* hops, u64* edges) {
    const size_t n = (size_t)ctx->g.n;
    const fora_params& p = ctx->p;
    int rc;
    CK(ctx->counts.ensure(n));
    CK(ctx->scratch64.ensure(4));
    CK(ctx->scratch32.ensure(4));
    CK(cudaMemsetAsync(ctx->counts.p, 0, sizeof(u64) * n, ctx->stream));
    CK(cudaMemsetAsync(ctx->scratch64.p, 0, sizeof(u64) * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->scratch32.p, 0, sizeof(int32_t) * 4, ctx->stream));
    const u64 nw = (u64)ceil(p.omega); // for(i=0; i<omega; i++), query.h:81
    u64 walk_hops = 0;
    if ((rc = launch_bulk_single(ctx, source, nw, nullptr, ctx->counts.p, 0x42500000u + qid, 0, &walk_hops))) return rc;
    if (p.rmax < 1.0) { // query.h:91
        if ((rc = ensure_bwd(ctx))) return rc;
        if ((rc = launch_bwd(ctx, source, p.rmax, p.omega, ctx->counts.p, d_ppr, 0, ctx->g.n, nullptr, 0, ctx->scratch64.p + 1, ctx->scratch32.p))) return rc;
    } else {            // query.h:114-119
        counts_to_ppr_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->g.n, ctx->counts.p, p.omega, d_ppr);
        CKL();
    }
    u64 h2[2];
    int32_t ovf = 0;
    CK(cudaMemcpyAsync(h2, ctx->scratch64.p, sizeof(u64) * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&ovf, ctx->scratch32.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ovf) { ctx->bwd_blocks = 0; return ctx->fail(FORA_ECUDA, "backward push: touched list overflow"); }
    *n_walks = nw; *hops = walk_hops; *edges = h2[1];
    return FORA_OK;
}


// =============================================================================================
// multi-GPU split of ONE query (Twitter-scale): every GPU holds the same push state
// (broadcast by the caller into fora_device_reserve / fora_device_residue of slot 0), walks its share of the
// walk chunks into its own dense vector and the caller sums the vectors (ncclAllReduce over NVLink).
// part 0 starts from the reserve (+ alpha*r credit with --opt), the other parts from zero.
// =============================================================================================
extern "C" void* fora_device_reserve(fora_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || slot >= ctx->alloc_slots) return nullptr;
    return ctx->reserve.p + (size_t)ctx->g.n * slot;
}
extern "C" void* fora_device_residue(fora_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || slot >= ctx->alloc_slots) return nullptr;
    return ctx->residue.p + (size_t)ctx->g.n * slot;
}
extern "C" int fora_prepare_slots(fora_ctx* ctx) { return require_ready(ctx, ctx ? ctx->p.omega : 0); }
// dense vector in the engine's internal vertex order (what the device pointers above hold) -> original ids
extern "C" int fora_device_to_original(fora_ctx* ctx, const double* d_internal, double* d_original) {
    if (!ctx || !ctx->g.n || !d_internal || !d_original || d_internal == d_original) return ctx ? ctx->fail(FORA_EINVAL, "bad arguments") : FORA_EINVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = vec_to_original(ctx, d_internal, d_original, (size_t)ctx->g.n);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return FORA_OK;
}

extern "C" int fora_compute_ppr_part_device(fora_ctx* ctx, double rsum, uint32_t qid, uint32_t part, uint32_t nparts, fora_query_stat* stat) {
    int rc = require_ready(ctx, ctx ? ctx->p.omega : 0);
    if (rc) return rc;
    if (nparts == 0 || part >= nparts) return ctx->fail(FORA_EINVAL, "bad part / nparts");
    const size_t n = (size_t)ctx->g.n;
    SlotMeta* h = ctx->h_meta;
    memset(h, 0, sizeof *h);
    for (int s = 0; s < MAX_SLOTS; ++s) h->source[s] = -1;
    h->source[0] = 0;
    h->qid[0] = qid;
    h->state[0] = rsum == 0.0 ? 2 : 1;
    h->rsum[0] = rsum;
    if ((rc = meta_h2d(ctx))) return rc;
    if (part > 0) CK(cudaMemsetAsync(ctx->reserve.p, 0, sizeof(double) * n, ctx->stream));
    if ((rc = walk_wave(ctx, ctx->reserve.p, 0, ctx->p.opt, ctx->p.opt, 0, nullptr, part, nparts))) return rc;
    if ((rc = meta_d2h_sync(ctx))) return rc;
    if (stat) fill_stat(ctx, 0, ctx->p.rmax, 0, stat);
    ctx->session_source = -1;
    return FORA_OK;
}

// =============================================================================================
// queries
// =============================================================================================
// Compacted result: (original id, value) of every entry >= threshold of the dense vectors of slots [slot_lo, slot_lo + gridDim.y),
// appended per slot (unordered).  A block counts first and reserves its places with one atomic.
__global__ void __launch_bounds__(256) sparse_out_kernel(int32_t n, const double* __restrict__ ppr, const int32_t* __restrict__ new2old, double thr,
                                                         u32 cap, int slot_lo, int32_t* __restrict__ ids, double* __restrict__ vals, u32* __restrict__ cnt) {
    __shared__ u32 s_cnt[8];
    __shared__ u32 s_base;
    const int slot = slot_lo + (int)blockIdx.y;
    const double* __restrict__ v = ppr + (size_t)slot * n;
    const int per_block = (n + gridDim.x - 1) / gridDim.x;
    const int lo = min(n, (int)blockIdx.x * per_block), hi = min(n, lo + per_block);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 mine = 0;
    for (int i = lo + (int)threadIdx.x; i < hi; i += 256) mine += v[i] >= thr;
    const u32 wi = warp_incl_scan(mine);
    if (lane == 31) s_cnt[w] = wi;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 acc = 0;
        for (int t = 0; t < 8; ++t) { const u32 x = s_cnt[t]; s_cnt[t] = acc; acc += x; }
        s_base = acc ? atomicAdd(&cnt[slot], acc) : 0u;
    }
    __syncthreads();
    u32 pos = s_base + s_cnt[w] + wi - mine;
    for (int i = lo + (int)threadIdx.x; i < hi; i += 256) {
        const double x = v[i];
        if (x >= thr) {
            if (pos < cap) {
                ids[(size_t)slot * cap + pos] = new2old ? new2old[i] : i;
                vals[(size_t)slot * cap + pos] = x;
            }
            ++pos;
        }
    }
}
struct SparseOut { // fora_query_batch_sparse
    double threshold;
    u64 cap_per_query, cap_total;
    int32_t* ids;
    double* vals;
    uint64_t* offsets; // [n_q + 1]
};

static int query_batch_impl(fora_ctx* ctx, int algo, const int32_t* h_sources, const int32_t* d_sources, int32_t n_q, double* ppr,
                            fora_query_stat* stats, fora_batch_timing* timing, const SparseOut* sp = nullptr) {
    int rc = require_ready(ctx, ctx ? ctx->p.omega : 0);
    if (rc) return rc;
    if (n_q < 0 || (!h_sources && !d_sources && n_q)) return ctx->fail(FORA_EINVAL, "bad sources");
    if (algo != FORA_ALGO_FORA && algo != FORA_ALGO_FWDPUSH && algo != FORA_ALGO_MC && algo != FORA_ALGO_BIPPR)
        return ctx->fail(FORA_EINVAL, "unknown algo");
    if (algo == FORA_ALGO_BIPPR && d_sources) return ctx->fail(FORA_EINVAL, "bippr: host sources only");
    const size_t n = (size_t)ctx->g.n;
    const int S = ctx->slots;
    ctx->session_source = -1;
    const u64 launches0 = ctx->launches;
    float push_ms = 0, walk_ms = 0, copy_ms = 0;
    ctx->push_kernel_ms = ctx->walk_kernel_ms = 0;
    ctx->push_kernel_launches = ctx->walk_kernel_launches = 0;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    std::vector<double> fr(S);
    std::vector<u64> rounds(S);
    SlotMeta* h = ctx->h_meta;
    u32 sp_cap = 0;
    u64 sp_total = 0;
    int sp_wave = 0;
    if (sp) { // compacted output: two device buffers (wave parity) so that the copy of wave w runs under wave w+1
        if (!sp->ids || !sp->vals || !sp->offsets || !(sp->threshold > 0.0)) return ctx->fail(FORA_EINVAL, "sparse output: ids / vals / offsets and a positive threshold are required");
        sp_cap = (u32)std::min<u64>({sp->cap_per_query, (u64)n, (u64)0xfffffff0u});
        if (sp_cap == 0) return ctx->fail(FORA_EINVAL, "sparse output: cap_per_query is 0");
        CK(ctx->sp_ids.ensure((size_t)2 * S * sp_cap));
        CK(ctx->sp_vals.ensure((size_t)2 * S * sp_cap));
        CK(ctx->sp_cnt.ensure((size_t)2 * MAX_SLOTS));
        sp->offsets[0] = 0;
        for (auto& e : ctx->ev_sp) if (!e) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    for (int32_t q0 = 0; q0 < n_q; q0 += S) {
        const int cnt = std::min<int32_t>(S, n_q - q0);
        for (int s = 0; s < cnt; ++s) {
            if (h_sources) {
                if (h_sources[q0 + s] < 0 || h_sources[q0 + s] >= ctx->g.n) return ctx->fail(FORA_EINVAL, "source out of range");
                h->source[s] = to_internal(ctx, h_sources[q0 + s]);
            } else h->source[s] = 0;
            h->qid[s] = (u32)(ctx->qid_base + (u64)(q0 + s));
        }
        CK(cudaEventRecord(ctx->ev[1], ctx->stream));
        // results leave through a staging buffer on a second stream, so the device->host copy of wave w overlaps the
        // computation of wave w+1 (the slot vectors are re-initialised immediately).  The caller has made sure the
        // previous wave's copy released the staging buffer.
        bool shipped = false;
        auto ship = [&](int lo, int hi) -> int {
            if (hi <= lo) return FORA_OK;
            for (int s = lo; s < hi; ++s) {
                int r2 = vec_to_original(ctx, ctx->reserve.p + n * s, ctx->stage.p + n * s, n); // back to original ids
                if (r2) return r2;
            }
            CK(cudaEventRecord(ctx->ev_stage_ready, ctx->stream));
            CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_stage_ready, 0));
            CK(cudaMemcpyAsync(ppr + ((size_t)q0 + lo) * n, ctx->stage.p + n * lo, sizeof(double) * n * (size_t)(hi - lo), cudaMemcpyDeviceToHost, ctx->copy_stream));
            CK(cudaEventRecord(ctx->ev_stage_free, ctx->copy_stream));
            ctx->stage_busy = true;
            return FORA_OK;
        };
        if (algo == FORA_ALGO_FORA) {
            if ((rc = push_wave(ctx, cnt, d_sources ? d_sources + q0 : nullptr, fr.data(), rounds.data()))) return rc;
            CK(cudaEventRecord(ctx->ev[2], ctx->stream));
            if (ppr && S >= 8) { // ship finished slot groups while the others still walk: the exposed copy tail shrinks 4x
                if (ctx->stage_busy) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_stage_free, 0));
                CK(ctx->stage.ensure(n * (size_t)S));
                if ((rc = walk_wave(ctx, ctx->reserve.p, 0, ctx->p.opt, ctx->p.opt, 0, nullptr, 0, 1, 4, [&](int lo, int hi) { return ship(lo, std::min(hi, cnt)); }))) return rc;
                shipped = true;
            } else if ((rc = walk_wave(ctx, ctx->reserve.p, 0, ctx->p.opt, ctx->p.opt, 0, nullptr))) return rc;
        } else if (algo == FORA_ALGO_FWDPUSH) { // query.h:1503-1508: push at config.rmax, ppr = reserve
            const int keep = ctx->p.balanced;
            ctx->p.balanced = 0;
            rc = push_wave(ctx, cnt, d_sources ? d_sources + q0 : nullptr, fr.data(), rounds.data());
            ctx->p.balanced = keep;
            if (rc) return rc;
            CK(cudaEventRecord(ctx->ev[2], ctx->stream));
        } else if (algo == FORA_ALGO_BIPPR) { // query.h:71-124
            CK(cudaEventRecord(ctx->ev[2], ctx->stream));
            for (int s = 0; s < MAX_SLOTS; ++s) { h->state[s] = 0; h->nwalk[s] = h->hops[s] = h->edges[s] = h->vertices[s] = h->levels[s] = h->nsrc[s] = h->idx_hits[s] = 0; h->rsum[s] = 0; }
            for (int s = 0; s < cnt; ++s) {
                u64 nw, hp, ed;
                if ((rc = bippr_one(ctx, h->source[s], h->qid[s], ctx->reserve.p + n * s, &nw, &hp, &ed))) return rc;
                h->nwalk[s] = nw; h->hops[s] = hp; h->edges[s] = ed; h->state[s] = 1;
                rounds[s] = 0;
            }
            if ((rc = meta_h2d(ctx))) return rc;
        } else { // Monte-Carlo, query.h:16-43: omega walks from the source, ppr = count/omega
            if ((rc = init_wave(ctx, cnt, 0, d_sources ? d_sources + q0 : nullptr))) return rc;
            if ((rc = meta_d2h_sync(ctx))) return rc;
            CK(cudaEventRecord(ctx->ev[2], ctx->stream));
            CK(ctx->counts.ensure(n));
            const u64 nw = (u64)ceil(ctx->p.omega); // for(i=0; i<omega; i++), query.h:25
            for (int s = 0; s < cnt; ++s) {
                CK(cudaMemsetAsync(ctx->counts.p, 0, sizeof(u64) * n, ctx->stream));
                int32_t src_int = h->source[s];
                if (d_sources) { // source id lives on the device: fetch it (tiny)
                    int32_t sv;
                    CK(cudaMemcpyAsync(&sv, d_sources + q0 + s, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
                    CK(cudaStreamSynchronize(ctx->stream));
                    src_int = to_internal(ctx, sv);
                }
                u64 hp = 0;
                if ((rc = launch_bulk_single(ctx, src_int, nw, nullptr, ctx->counts.p, 0x4d430000u + h->qid[s], 0, &hp))) return rc;
                h->hops[s] = hp;
                counts_to_ppr_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->g.n, ctx->counts.p, ctx->p.omega, ctx->reserve.p + n * s);
                CKL();
                h->nwalk[s] = nw;
            }
            CK(cudaMemcpyAsync(ctx->meta.p->nwalk, h->nwalk, sizeof(u64) * MAX_SLOTS, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->meta.p->hops, h->hops, sizeof(u64) * MAX_SLOTS, cudaMemcpyHostToDevice, ctx->stream));
        }
        CK(cudaEventRecord(ctx->ev[3], ctx->stream));
        u32* sp_cnt_dev = nullptr;
        if (sp) { // compaction on the work stream; counts reach the host with the wave's control block below
            const int b = sp_wave & 1;
            if (ctx->sp_busy[b]) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_sp[b], 0)); // the copy of wave w-2 has left this buffer
            sp_cnt_dev = ctx->sp_cnt.p + (size_t)b * MAX_SLOTS;
            CK(cudaMemsetAsync(sp_cnt_dev, 0, sizeof(u32) * MAX_SLOTS, ctx->stream));
            const int gx = std::max(1, std::min(ctx->num_sms * 8 / std::max(cnt, 1) + 1, (int)((n + 2047) / 2048)));
            sparse_out_kernel<<<dim3(gx, cnt), 256, 0, ctx->stream>>>(ctx->g.n, ctx->reserve.p, ctx->g.relabeled ? ctx->g.new2old : nullptr, sp->threshold, sp_cap, 0,
                                                                   ctx->sp_ids.p + (size_t)b * S * sp_cap, ctx->sp_vals.p + (size_t)b * S * sp_cap, sp_cnt_dev);
            CKL();
            CK(cudaMemcpyAsync(ctx->meta.p->sp_count, sp_cnt_dev, sizeof(u32) * MAX_SLOTS, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        if (ppr && !shipped) {
            CK(ctx->stage.ensure(n * (size_t)S));
            if (ctx->stage_busy) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_stage_free, 0));
            if ((rc = ship(0, cnt))) return rc;
        }
        CK(cudaEventRecord(ctx->ev[4], ctx->stream));
        if ((rc = meta_d2h_sync(ctx))) return rc;
        if (sp) { // the wave has finished (sync above): ship exactly the entries found, on the copy stream, under the next wave
            const int b = sp_wave & 1;
            for (int s = 0; s < cnt; ++s) {
                const u64 c = h->sp_count[s];
                if (c > sp_cap) return ctx->fail(FORA_ERANGE, "sparse output: a query has more entries >= threshold than cap_per_query");
                if (sp_total + c > sp->cap_total) return ctx->fail(FORA_ERANGE, "sparse output: cap_total exceeded");
                if (c) {
                    CK(cudaMemcpyAsync(sp->ids + sp_total, ctx->sp_ids.p + ((size_t)b * S + s) * sp_cap, sizeof(int32_t) * c, cudaMemcpyDeviceToHost, ctx->copy_stream));
                    CK(cudaMemcpyAsync(sp->vals + sp_total, ctx->sp_vals.p + ((size_t)b * S + s) * sp_cap, sizeof(double) * c, cudaMemcpyDeviceToHost, ctx->copy_stream));
                }
                sp_total += c;
                sp->offsets[q0 + s + 1] = sp_total;
            }
            CK(cudaEventRecord(ctx->ev_sp[b], ctx->copy_stream));
            ctx->sp_busy[b] = true;
            ++sp_wave;
        }
        if (stats)
            for (int s = 0; s < cnt; ++s) fill_stat(ctx, s, algo == FORA_ALGO_FORA ? fr[s] : ctx->p.rmax, algo == FORA_ALGO_MC ? 0 : rounds[s], &stats[q0 + s]);
        float t;
        CK(cudaEventElapsedTime(&t, ctx->ev[1], ctx->ev[2])); push_ms += t;
        CK(cudaEventElapsedTime(&t, ctx->ev[2], ctx->ev[3])); walk_ms += t;
        CK(cudaEventElapsedTime(&t, ctx->ev[3], ctx->ev[4])); copy_ms += t;
    }
    if (ctx->stage_busy) { // the last copy must have landed before the call returns
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_stage_free, 0));
        ctx->stage_busy = false;
    }
    for (int b = 0; b < 2; ++b)
        if (ctx->sp_busy[b]) {
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_sp[b], 0));
            ctx->sp_busy[b] = false;
        }
    CK(cudaEventRecord(ctx->ev[5], ctx->stream));
    CK(cudaEventSynchronize(ctx->ev[5]));
    if (timing) {
        memset(timing, 0, sizeof *timing);
        CK(cudaEventElapsedTime(&timing->total_ms, ctx->ev[0], ctx->ev[5]));
        timing->push_ms = push_ms; timing->walk_ms = walk_ms; timing->copy_ms = copy_ms;
        timing->kernel_launches = ctx->launches - launches0;
        timing->push_kernel_ms = (float)ctx->push_kernel_ms; timing->walk_kernel_ms = (float)ctx->walk_kernel_ms;
        timing->push_kernel_launches = ctx->push_kernel_launches; timing->walk_kernel_launches = ctx->walk_kernel_launches;
    }
    return FORA_OK;
}

extern "C" int fora_query_batch(fora_ctx* ctx, int algo, const int32_t* sources, int32_t n_q, double* ppr,
                                fora_query_stat* stats, fora_batch_timing* timing) {
    return query_batch_impl(ctx, algo, sources, nullptr, n_q, ppr, stats, timing);
}
extern "C" int fora_query_batch_sparse(fora_ctx* ctx, int algo, const int32_t* sources, int32_t n_q, double threshold, uint64_t cap_per_query,
                                       uint64_t cap_total, int32_t* ids, double* values, uint64_t* offsets, fora_query_stat* stats,
                                       fora_batch_timing* timing) {
    if (!ctx) return FORA_EINVAL;
    SparseOut sp{threshold, cap_per_query, cap_total, ids, values, offsets};
    return query_batch_impl(ctx, algo, sources, nullptr, n_q, nullptr, stats, timing, &sp);
}
extern "C" int fora_query_batch_device(fora_ctx* ctx, int algo, const int32_t* d_sources, int32_t n_q, fora_query_stat* stats,
                                       fora_batch_timing* timing) {
    return query_batch_impl(ctx, algo, nullptr, d_sources, n_q, nullptr, stats, timing);
}
extern "C" void* fora_device_ppr(fora_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || slot >= ctx->alloc_slots) return nullptr;
    return ctx->reserve.p + (size_t)ctx->g.n * slot;
}


// =============================================================================================
// top-k queries: get_topk(), query.h:1139-1190
// =============================================================================================
// advance the per-source index cursor after a round: rw_counter[source] += min(n_v, remaining) (query.h:588,603)
__global__ void idx_cursor_kernel(int32_t n, const int32_t* __restrict__ srcs, const u64* __restrict__ woff,
                                  const u64* __restrict__ nsrc, const u64* __restrict__ idx_cnt, u64* __restrict__ idx_used,
                                  const int32_t* __restrict__ slot_state) {
    const int slot = blockIdx.y;
    if (slot_state[slot] != 1) return;
    const u64 ns = nsrc[slot];
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < ns; i += (u64)gridDim.x * blockDim.x) {
        const int32_t v = srcs[(size_t)slot * n + i];
        const u64 n_v = woff[(size_t)slot * (n + 1) + i + 1] - woff[(size_t)slot * (n + 1) + i];
        u64* u = &idx_used[(size_t)slot * n + v];
        const u64 remaining = idx_cnt[v] - *u;
        *u += n_v < remaining ? n_v : remaining;
    }
}

extern "C" int fora_topk_batch(fora_ctx* ctx, int algo, const int32_t* sources, int32_t n_q, uint32_t k, int32_t* nodes,
                               double* values, int32_t* iters, fora_query_stat* stats, fora_batch_timing* timing) {
    if (!ctx) return FORA_EINVAL;
    if (!ctx->params_set) return ctx->fail(FORA_EINVAL, "fora_params_set has not been called");
    if (!ctx->g.n) return ctx->fail(FORA_EINVAL, "no graph uploaded");
    const int32_t n = ctx->g.n;
    if (!(k > 1 && (int64_t)k < (int64_t)n - 1)) return ctx->fail(FORA_EINVAL, "k must satisfy 1 < k < n-1 (query.h:1317-1318)");
    if (!sources || !nodes || !values || n_q < 0) return ctx->fail(FORA_EINVAL, "bad arguments");
    const fora_params keep = ctx->p;
    int rc = FORA_OK;
    auto restore = [&](int code) { ctx->p = keep; return code; };
    const double min_delta = 1.0 / n;
    if (algo == FORA_ALGO_FORA) {
        // the largest omega of either driver is reached at delta = 1/n with its own pfail
        // (fora_query_topk_new: 1/n^2, query.h:977; fora_query_topk_with_bound: 1/n^2/ln n, query.h:915)
        double rm, om;
        fora_host_setting(keep.opt ? 1 : 0, n, ctx->g.m_decl, keep.epsilon, min_delta, keep.opt ? 1.0 / n / n : 1.0 / n / n / log((double)n), keep.alpha,
                          keep.opt, keep.rmax_scale, &rm, &om);
        if ((rc = require_ready(ctx, om))) return rc;
    } else if ((rc = require_ready(ctx, keep.omega))) return rc;
    const int S = ctx->slots;
    const size_t nn = (size_t)n;
    ctx->session_source = -1;
    const u64 launches0 = ctx->launches;
    ctx->push_kernel_ms = ctx->walk_kernel_ms = 0;
    ctx->push_kernel_launches = ctx->walk_kernel_launches = 0;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    float topk_ms = 0;
    SlotMeta* h = ctx->h_meta;
    std::vector<double> fr(S);
    std::vector<u64> rounds(S);
    for (int32_t q0 = 0; q0 < n_q; q0 += S) {
        const int cnt = std::min<int32_t>(S, n_q - q0);
        for (int s = 0; s < cnt; ++s) {
            if (sources[q0 + s] < 0 || sources[q0 + s] >= n) return restore(ctx->fail(FORA_EINVAL, "source out of range"));
            h->source[s] = to_internal(ctx, sources[q0 + s]);
            h->qid[s] = (u32)(ctx->qid_base + (u64)(q0 + s));
        }
        const double* result = ctx->reserve.p; // where each slot's final vector lives
        std::vector<int32_t> it(cnt, 0);
        std::vector<u64> tot_walks(cnt, 0), tot_hits(cnt, 0), tot_hops(cnt, 0);
        if (algo == FORA_ALGO_FORA && !keep.opt) {
            // ---- fora_query_topk_with_bound (query.h:909-969): delta from 1/4 halving, per-node bounds, if_stop ----
            CK(ctx->ppr.ensure(nn * S));
            CK(ctx->ub.ensure(nn * S));
            CK(ctx->lb.ensure(nn * S));
            CK(ctx->in_topk.ensure(nn * S)); // one byte map per slot: the bound tests of a round run for all slots at once
            CK(ctx->flags.ensure(2 * MAX_SLOTS));
            const bool use_idx = keep.with_idx && ctx->has_index;
            if (use_idx) {
                CK(ctx->idx_used.ensure(nn * S));
                CK(cudaMemsetAsync(ctx->idx_used.p, 0, sizeof(u64) * nn * cnt, ctx->stream)); // query.h:935-936
            }
            if ((rc = init_wave(ctx, cnt, 0, nullptr))) return restore(rc);
            if ((rc = meta_d2h_sync(ctx))) return restore(rc);
            std::vector<char> done(cnt, 0);
            for (int s = 0; s < cnt; ++s) {
                done[s] = h->state[s] != 1;
                if (h->state[s] == 2) it[s] = 1;
            }
            fill_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->ub.p, nn * cnt, 1.0); // upper_bounds.reset_one_values(), query.h:940
            CKL();
            CK(cudaMemsetAsync(ctx->lb.p, 0, sizeof(double) * nn * cnt, ctx->stream));          // lower_bounds.reset_zero_values()
            CK(cudaMemsetAsync(ctx->in_topk.p, 0, nn * cnt, ctx->stream));
            CK(cudaMemcpyAsync(ctx->ppr.p, ctx->reserve.p, sizeof(double) * nn * cnt, cudaMemcpyDeviceToDevice, ctx->stream));
            const double pfail = 1.0 / n / n / log((double)n);                               // query.h:915
            const double threshold = (1.0 - 0.77) / pow(500, 0.77) / pow((double)n, 1 - 0.77); // query.h:913, ppr_decay_alpha = 0.77
            double delta = 1.0 / 4;                                                           // query.h:912
            for (int round = 0; round < 64 && delta >= min_delta; ++round) {
                bool any = false;
                double rmax, omega;
                fora_host_setting(0, n, ctx->g.m_decl, keep.epsilon, delta, pfail, keep.alpha, 0, keep.rmax_scale, &rmax, &omega); // fora_setting, query.h:944
                for (int s = 0; s < MAX_SLOTS; ++s) h->active[s] = 0;
                for (int s = 0; s < cnt; ++s)
                    if (!done[s]) { h->active[s] = 1; h->rmax[s] = rmax; any = true; it[s]++; }
                if (!any) break;
                if ((rc = push_round_active(ctx))) return restore(rc);
                for (int s = 0; s < cnt; ++s) h->state[s] = done[s] ? (h->state[s] == 1 ? 3 : h->state[s]) : 1;
                CK(cudaMemcpyAsync(ctx->meta.p->state, h->state, sizeof(int32_t) * MAX_SLOTS, cudaMemcpyHostToDevice, ctx->stream));
                for (int s = 0; s < cnt; ++s)
                    if (!done[s]) CK(cudaMemcpyAsync(ctx->ppr.p + nn * s, ctx->reserve.p + nn * s, sizeof(double) * nn, cudaMemcpyDeviceToDevice, ctx->stream));
                ctx->p.omega = omega;
                ctx->p.rmax = rmax;
                // walks: plain random_walk, n_v = ceil(r/rsum*N) (no index, query.h:724-740) or ceil(r*omega) (index, 656-662)
                if ((rc = walk_wave(ctx, ctx->ppr.p, use_idx ? 3 : 0, 0, 0, (u32)(round + 1), use_idx ? ctx->idx_used.p : nullptr))) return restore(rc);
                if (use_idx) {
                    idx_cursor_kernel<<<dim3(ctx->num_sms * 2, S), 256, 0, ctx->stream>>>(n, ctx->srcs.p, ctx->woff.p, ctx->meta.p->nsrc, ctx->idx_cnt.p, ctx->idx_used.p, ctx->meta.p->state);
                    CKL();
                }
                if ((rc = meta_d2h_sync(ctx))) return restore(rc);
                CK(cudaEventRecord(ctx->ev[6], ctx->stream));
                std::vector<int> act, need_lb;
                for (int s = 0; s < cnt; ++s) {
                    if (done[s]) continue;
                    tot_walks[s] += h->nwalk[s]; tot_hits[s] += h->idx_hits[s]; tot_hops[s] += h->hops[s];
                    act.push_back(s);
                }
                // set_ppr_bounds of every active slot in one launch (query.h:745-746; the per-slot scalars are read on the device)
                CK(ctx->sel_slots.ensure(MAX_SLOTS));
                CK(ctx->stop_slots.ensure(MAX_SLOTS));
                CK(ctx->stop_lowk.ensure(MAX_SLOTS));
                CK(ctx->flags.ensure(2 * MAX_SLOTS));
                if (!act.empty()) {
                    int32_t ids[MAX_SLOTS];
                    for (size_t j = 0; j < act.size(); ++j) ids[j] = act[j];
                    CK(cudaMemcpyAsync(ctx->stop_slots.p, ids, sizeof(int32_t) * act.size(), cudaMemcpyHostToDevice, ctx->stream));
                    CK(cudaStreamSynchronize(ctx->stream)); // ids lives on this stack frame
                    const int gx = std::max(1, std::min(ctx->num_sms * 8 / (int)act.size() + 1, (n + 255) / 256));
                    ppr_bounds_kernel<<<dim3(gx, (unsigned)act.size()), 256, 0, ctx->stream>>>(n, nn, ctx->stop_slots.p, ctx->meta.p->rsum, ctx->meta.p->nwalk, pfail, delta, threshold,
                                                                                            ctx->ppr.p, ctx->reserve.p, ctx->ub.p, ctx->lb.p);
                    CKL();
                }
                // if_stop(), algo.h:1096-1166: k-th estimate of every active slot in one batched select ...
                std::vector<SelResult> sres(MAX_SLOTS);
                std::vector<char> stop(cnt, 0);
                if ((rc = select_batch(ctx, ctx->ppr.p, nn, act.data(), (int)act.size(), k, false, sres.data()))) return restore(rc);
                for (size_t j = 0; j < act.size(); ++j) {
                    if (sel_kth(sres[j]) >= 2.0 * delta) stop[act[j]] = 1;
                    else if (!(delta >= threshold)) need_lb.push_back(act[j]);
                }
                // ... then the k largest lower bounds of the slots that are still open, and the bound tests of all of them in three launches
                if ((rc = select_batch(ctx, ctx->lb.p, nn, need_lb.data(), (int)need_lb.size(), k, false, sres.data()))) return restore(rc);
                {
                    // only slots with k positive lower bounds above delta can pass (fewer: some top-k node has lower bound 0, algo.h:1129-1134)
                    std::vector<int> test_j;
                    for (size_t j = 0; j < need_lb.size(); ++j)
                        if (sres[j].out_count == k && sel_kth(sres[j]) > delta) test_j.push_back((int)j);
                    if (!test_j.empty()) {
                        // the select left vector j's nodes at sel_on + j*sel_p2; the stop kernels index lists by position in need_lb
                        int32_t ids[MAX_SLOTS];
                        double lowk[MAX_SLOTS];
                        for (size_t j = 0; j < need_lb.size(); ++j) { ids[j] = need_lb[j]; lowk[j] = sel_kth(sres[j]); }
                        CK(cudaMemcpyAsync(ctx->stop_slots.p, ids, sizeof(int32_t) * need_lb.size(), cudaMemcpyHostToDevice, ctx->stream));
                        CK(cudaMemcpyAsync(ctx->stop_lowk.p, lowk, sizeof(double) * need_lb.size(), cudaMemcpyHostToDevice, ctx->stream));
                        CK(cudaMemsetAsync(ctx->flags.p, 0, sizeof(u32) * 2 * MAX_SLOTS, ctx->stream));
                        const unsigned nj = (unsigned)need_lb.size();
                        const int gx = std::max(1, std::min(ctx->num_sms * 8 / (int)nj + 1, (n + 255) / 256));
                        stop_mark_kernel<<<dim3(4, nj), 256, 0, ctx->stream>>>(ctx->sel_on.p, ctx->sel_p2, k, ctx->stop_slots.p, nn, ctx->ub.p, ctx->lb.p, keep.epsilon, ctx->in_topk.p, ctx->flags.p);
                        stop_tail_kernel<<<dim3(gx, nj), 256, 0, ctx->stream>>>(n, ctx->stop_slots.p, nn, ctx->ppr.p, ctx->ub.p, ctx->lb.p, ctx->in_topk.p, ctx->stop_lowk.p, keep.epsilon, ctx->flags.p);
                        stop_unmark_kernel<<<dim3(4, nj), 256, 0, ctx->stream>>>(ctx->sel_on.p, ctx->sel_p2, k, ctx->stop_slots.p, nn, ctx->in_topk.p);
                        ctx->launches += 3;
                        u32 hf[2 * MAX_SLOTS];
                        CK(cudaMemcpyAsync(hf, ctx->flags.p, sizeof(u32) * 2 * nj, cudaMemcpyDeviceToHost, ctx->stream));
                        CK(cudaStreamSynchronize(ctx->stream));
                        for (int j : test_j) stop[need_lb[(size_t)j]] = !hf[2 * j] && !hf[2 * j + 1];
                    }
                }
                for (int s : act)
                    if (stop[s] || delta <= min_delta) done[s] = 1; // query.h:963
                CK(cudaEventRecord(ctx->ev[7], ctx->stream));
                CK(cudaEventSynchronize(ctx->ev[7]));
                float t;
                CK(cudaEventElapsedTime(&t, ctx->ev[6], ctx->ev[7]));
                topk_ms += t;
                CK(cudaMemsetAsync(ctx->meta.p->hops, 0, sizeof(u64) * MAX_SLOTS, ctx->stream));
                CK(cudaMemsetAsync(ctx->meta.p->idx_hits, 0, sizeof(u64) * MAX_SLOTS, ctx->stream));
                delta = std::max(min_delta, delta / 2.0); // query.h:966
                bool all = true;
                for (int s = 0; s < cnt; ++s) all = all && done[s];
                if (all) break;
            }
            result = ctx->ppr.p;
            for (int s = 0; s < cnt; ++s) { fr[s] = ctx->p.rmax; rounds[s] = (u64)it[s]; }
            ctx->p = keep;
        } else if (algo == FORA_ALGO_FORA) {
            CK(ctx->ppr.ensure(nn * S));
            const bool use_idx = keep.with_idx && ctx->has_index;
            if (use_idx) {
                CK(ctx->idx_used.ensure(nn * S));
                CK(cudaMemsetAsync(ctx->idx_used.p, 0, sizeof(u64) * nn * cnt, ctx->stream)); // rw_counter.reset_zero_values(), query.h:998
            }
            if ((rc = init_wave(ctx, cnt, 0, nullptr))) return restore(rc);
            if ((rc = meta_d2h_sync(ctx))) return restore(rc);
            std::vector<char> done(cnt, 0);
            for (int s = 0; s < cnt; ++s) {
                done[s] = h->state[s] != 1; // source without out-edges: ppr = {s:1} after one iteration (query.h:1003-1011)
                if (h->state[s] == 2) it[s] = 1;
            }
            double delta = 1.0 / k / 10;                               // query.h:976
            const double pfail = 1.0 / n / n;                          // query.h:977
            CK(cudaMemcpyAsync(ctx->ppr.p, ctx->reserve.p, sizeof(double) * nn * cnt, cudaMemcpyDeviceToDevice, ctx->stream));
            for (int round = 0; round < 64 && delta >= min_delta; ++round) {
                bool any = false;
                double rmax, omega;
                fora_host_setting(1, n, ctx->g.m_decl, keep.epsilon, delta, pfail, keep.alpha, keep.opt, keep.rmax_scale, &rmax, &omega);
                for (int s = 0; s < MAX_SLOTS; ++s) h->active[s] = 0;
                for (int s = 0; s < cnt; ++s)
                    if (!done[s]) { h->active[s] = 1; h->rmax[s] = rmax; any = true; it[s]++; }
                if (!any) break;
                if ((rc = push_round_active(ctx))) return restore(rc);
                // finished slots keep their vector: only active slots are rebuilt and walked this round
                for (int s = 0; s < cnt; ++s) h->state[s] = done[s] ? (h->state[s] == 1 ? 3 : h->state[s]) : 1;
                CK(cudaMemcpyAsync(ctx->meta.p->state, h->state, sizeof(int32_t) * MAX_SLOTS, cudaMemcpyHostToDevice, ctx->stream));
                for (int s = 0; s < cnt; ++s)
                    if (!done[s]) CK(cudaMemcpyAsync(ctx->ppr.p + nn * s, ctx->reserve.p + nn * s, sizeof(double) * nn, cudaMemcpyDeviceToDevice, ctx->stream));
                ctx->p.omega = omega;
                ctx->p.rmax = rmax;
                // with index: alpha*r credit, (1-alpha) scaling, no-zero-hop walks (query.h:555-613); without: plain walks
                // of ceil(r*omega) each (query.h:615-632)
                if ((rc = walk_wave(ctx, ctx->ppr.p, 1, use_idx ? 1 : 0, use_idx ? 1 : 0, (u32)(round + 1), use_idx ? ctx->idx_used.p : nullptr))) return restore(rc);
                if (use_idx) {
                    idx_cursor_kernel<<<dim3(ctx->num_sms * 2, S), 256, 0, ctx->stream>>>(n, ctx->srcs.p, ctx->woff.p, ctx->meta.p->nsrc, ctx->idx_cnt.p, ctx->idx_used.p, ctx->meta.p->state);
                    CKL();
                }
                if ((rc = meta_d2h_sync(ctx))) return restore(rc);
                CK(cudaEventRecord(ctx->ev[6], ctx->stream));
                std::vector<int> act;
                for (int s = 0; s < cnt; ++s) {
                    if (done[s]) continue;
                    tot_walks[s] += h->nwalk[s]; tot_hits[s] += h->idx_hits[s]; tot_hops[s] += h->hops[s];
                    act.push_back(s);
                }
                std::vector<SelResult> sres(MAX_SLOTS);
                if ((rc = select_batch(ctx, ctx->ppr.p, nn, act.data(), (int)act.size(), k, false, sres.data()))) return restore(rc); // kth_ppr(), algo.h:578-590
                for (size_t j = 0; j < act.size(); ++j)
                    if (sel_kth(sres[j]) >= (1 + keep.epsilon) * delta || delta <= min_delta) done[act[j]] = 1; // query.h:1029
                CK(cudaEventRecord(ctx->ev[7], ctx->stream));
                CK(cudaEventSynchronize(ctx->ev[7]));
                float t;
                CK(cudaEventElapsedTime(&t, ctx->ev[6], ctx->ev[7]));
                topk_ms += t;
                CK(cudaMemsetAsync(ctx->meta.p->hops, 0, sizeof(u64) * MAX_SLOTS, ctx->stream));
                CK(cudaMemsetAsync(ctx->meta.p->idx_hits, 0, sizeof(u64) * MAX_SLOTS, ctx->stream));
                delta = std::max(min_delta, delta / 4.0); // query.h:1041
                bool all = true;
                for (int s = 0; s < cnt; ++s) all = all && done[s];
                if (all) break;
            }
            result = ctx->ppr.p;
            for (int s = 0; s < cnt; ++s) { fr[s] = ctx->p.rmax; rounds[s] = (u64)it[s]; }
            ctx->p = keep;
        } else {
            // fwdpush / montecarlo: the plain query, then topk_ppr (query.h:1141-1167)
            std::vector<fora_query_stat> st(cnt);
            if ((rc = query_batch_impl(ctx, algo, sources + q0, nullptr, cnt, nullptr, st.data(), nullptr))) return restore(rc);
            if (stats) for (int s = 0; s < cnt; ++s) stats[q0 + s] = st[s];
        }
        CK(cudaEventRecord(ctx->ev[6], ctx->stream));
        { // topk_ppr(), algo.h:592-610: all slots of the wave in one batched select, sorted on the device
            std::vector<int> all(cnt);
            for (int s = 0; s < cnt; ++s) all[s] = s;
            std::vector<SelResult> sres(MAX_SLOTS);
            if ((rc = select_batch(ctx, result, nn, all.data(), cnt, k, true, sres.data()))) return restore(rc);
            for (int s = 0; s < cnt; ++s) {
                const u32 got = sres[s].out_count;
                if (got) {
                    CK(cudaMemcpyAsync(nodes + (size_t)(q0 + s) * k, ctx->sel_on.p + (size_t)ctx->sel_p2 * s, sizeof(int32_t) * got, cudaMemcpyDeviceToHost, ctx->stream));
                    CK(cudaMemcpyAsync(values + (size_t)(q0 + s) * k, ctx->sel_ov.p + (size_t)ctx->sel_p2 * s, sizeof(double) * got, cudaMemcpyDeviceToHost, ctx->stream));
                }
            }
            CK(cudaStreamSynchronize(ctx->stream));
            for (int s = 0; s < cnt; ++s) {
                int32_t* nd = nodes + (size_t)(q0 + s) * k;
                double* vl = values + (size_t)(q0 + s) * k;
                for (uint32_t j = sres[s].out_count; j < k; ++j) { nd[j] = 0; vl[j] = 0.0; } // unfilled slots stay (0, 0.0)
                if (ctx->g.relabeled) // internal -> original ids
                    for (uint32_t j = 0; j < sres[s].out_count; ++j) nd[j] = to_original(ctx, nd[j]);
            }
        }
        CK(cudaEventRecord(ctx->ev[7], ctx->stream));
        CK(cudaEventSynchronize(ctx->ev[7]));
        float t;
        CK(cudaEventElapsedTime(&t, ctx->ev[6], ctx->ev[7]));
        topk_ms += t;
        if (algo == FORA_ALGO_FORA) {
            if ((rc = meta_d2h_sync(ctx))) return restore(rc);
            for (int s = 0; s < cnt; ++s) {
                if (iters) iters[q0 + s] = it[s];
                if (stats) {
                    fill_stat(ctx, s, fr[s], rounds[s], &stats[q0 + s]);
                    stats[q0 + s].n_walks = tot_walks[s]; stats[q0 + s].n_idx_hits = tot_hits[s]; stats[q0 + s].walk_hops = tot_hops[s];
                }
            }
        } else if (iters) {
            for (int s = 0; s < cnt; ++s) iters[q0 + s] = 1;
        }
    }
    CK(cudaEventRecord(ctx->ev[5], ctx->stream));
    CK(cudaEventSynchronize(ctx->ev[5]));
    if (timing) {
        memset(timing, 0, sizeof *timing);
        CK(cudaEventElapsedTime(&timing->total_ms, ctx->ev[0], ctx->ev[5]));
        timing->topk_ms = topk_ms;
        timing->kernel_launches = ctx->launches - launches0;
        timing->push_kernel_ms = (float)ctx->push_kernel_ms; timing->walk_kernel_ms = (float)ctx->walk_kernel_ms;
        timing->push_kernel_launches = ctx->push_kernel_launches; timing->walk_kernel_launches = ctx->walk_kernel_launches;
    }
    return restore(FORA_OK);
}

// =============================================================================================
// walk index
// =============================================================================================
extern "C" int fora_index_info(fora_ctx* ctx, uint64_t* offsets, uint64_t* counts, uint64_t* total) {
    if (!ctx || !ctx->g.n) return ctx ? ctx->fail(FORA_EINVAL, "no graph") : FORA_EINVAL;
    if (!ctx->params_set || !offsets || !counts) return ctx->fail(FORA_EINVAL, "params / outputs missing");
    CK(cudaSetDevice(ctx->device));
    const int32_t n = ctx->g.n;
    std::vector<int32_t> deg_int((size_t)n), deg((size_t)n);
    CK(cudaMemcpy(deg_int.data(), ctx->g.deg, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost));
    for (int32_t v = 0; v < n; ++v) deg[(size_t)v] = deg_int[(size_t)to_internal(ctx, v)]; // offsets / counts are per ORIGINAL vertex id
    // build.h:325-334 -- host arithmetic in the reference's expression order (bit-exact counts)
    const fora_params& p = ctx->p;
    u64 tuned = 0;
    for (int32_t v = 0; v < n; ++v) {
        const size_t d = (size_t)deg[v];
        unsigned long num_rw;
        if (p.opt) num_rw = (unsigned long)ceil(d * p.rmax * (1 - p.alpha) * p.omega);
        else num_rw = (unsigned long)ceil(d * p.rmax * p.omega);
        offsets[v] = tuned;
        counts[v] = num_rw;
        tuned += num_rw;
    }
    if (total) *total = tuned;
    return FORA_OK;
}

extern "C" int fora_index_build(fora_ctx* ctx, const uint64_t* offsets, const uint64_t* counts, int32_t v_begin, int32_t v_end,
                                int32_t* dest) {
    if (!ctx || !ctx->g.n) return ctx ? ctx->fail(FORA_EINVAL, "no graph") : FORA_EINVAL;
    if (!offsets || !counts || !dest || v_begin < 0 || v_end > ctx->g.n || v_begin > v_end) return ctx->fail(FORA_EINVAL, "bad range");
    CK(cudaSetDevice(ctx->device));
    const u64 nseg = (u64)(v_end - v_begin);
    ctx->bulk_walks = ctx->bulk_hops = 0;
    ctx->bulk_kernel_ms = 0;
    if (nseg == 0) return FORA_OK;
    const u64 base = offsets[v_begin];
    const u64 total = offsets[v_end - 1] + counts[v_end - 1] - base;
    if (total == 0) return FORA_OK;
    // the walk plan of the query path: sources with at least one walk (internal ids, in ORIGINAL vertex order = index order)
    // and the exclusive prefix of their counts, which is the index offset relative to this source range (build.h:337-352)
    std::vector<int32_t> srcs;
    std::vector<u64> woff;
    srcs.reserve(nseg);
    woff.reserve(nseg + 1);
    for (u64 i = 0; i < nseg; ++i) {
        if (counts[v_begin + i] == 0) continue;
        if (offsets[v_begin + i] - base + counts[v_begin + i] > total) return ctx->fail(FORA_EINVAL, "index info is not an exclusive prefix of the counts");
        srcs.push_back(to_internal(ctx, (int32_t)(v_begin + (int64_t)i)));
        woff.push_back(offsets[v_begin + i] - base);
    }
    woff.push_back(total);
    const u64 ns = srcs.size();
    CK(ctx->scratch64.ensure(ns + 1));
    CK(ctx->scratch32.ensure(total + ns));
    int32_t* d_srcs = ctx->scratch32.p + total;
    CK(cudaMemcpyAsync(ctx->scratch64.p, woff.data(), sizeof(u64) * (ns + 1), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_srcs, srcs.data(), sizeof(int32_t) * ns, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream)); // the host vectors go out of scope below
    BulkPlan bp;
    bp.d_srcs = d_srcs; bp.d_woff = ctx->scratch64.p; bp.nsrc = ns; bp.nwalk = total; bp.out_dest = ctx->scratch32.p;
    bp.key_tag = 0x1d800000u; bp.no_zero_hop = ctx->p.opt; // build.h:347-350
    // destinations leave in slices while the remaining walks still run: slice p is copied on the copy stream behind launch p
    const int parts = (int)std::max<u64>(1, std::min<u64>(16, total / (8u << 20)));
    std::vector<cudaEvent_t> evs((size_t)parts);
    for (auto& e : evs) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    const double ms0 = ctx->walk_kernel_ms;
    u64 hops = 0;
    int rc = launch_bulk(ctx, bp, &hops, parts, [&](int part, u64 w_lo, u64 w_hi) -> int {
        CK(cudaEventRecord(evs[(size_t)part], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->copy_stream, evs[(size_t)part], 0));
        if (w_hi > w_lo) CK(cudaMemcpyAsync(dest + w_lo, ctx->scratch32.p + w_lo, sizeof(int32_t) * (w_hi - w_lo), cudaMemcpyDeviceToHost, ctx->copy_stream));
        return FORA_OK;
    });
    cudaError_t ce = cudaStreamSynchronize(ctx->copy_stream);
    for (auto& e : evs) cudaEventDestroy(e);
    if (rc) return rc;
    if (ce != cudaSuccess) return ctx->fail(FORA_ECUDA, std::string("index copy: ") + cudaGetErrorString(ce));
    ctx->bulk_walks = total;
    ctx->bulk_hops = hops;
    ctx->bulk_kernel_ms = ctx->walk_kernel_ms - ms0;
    return FORA_OK;
}
// walks, hops and walk-kernel milliseconds (CUDA events on the work stream) of the last fora_index_build call: roofline accounting
extern "C" int fora_index_build_stat(fora_ctx* ctx, uint64_t* walks, uint64_t* hops, double* kernel_ms) {
    if (!ctx) return FORA_EINVAL;
    if (walks) *walks = ctx->bulk_walks;
    if (hops) *hops = ctx->bulk_hops;
    if (kernel_ms) *kernel_ms = ctx->bulk_kernel_ms;
    return FORA_OK;
}

extern "C" int fora_index_upload(fora_ctx* ctx, const uint64_t* offsets, const uint64_t* counts, const int32_t* dest, uint64_t len) {
    if (!ctx || !ctx->g.n) return ctx ? ctx->fail(FORA_EINVAL, "no graph") : FORA_EINVAL;
    if (!offsets || !counts || (!dest && len)) return ctx->fail(FORA_EINVAL, "null index arrays");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->g.n;
    if (n && offsets[n - 1] + counts[n - 1] > len) return ctx->fail(FORA_EINVAL, "index info exceeds destination array");
    CK(ctx->idx_off.ensure(n));
    CK(ctx->idx_cnt.ensure(n));
    CK(ctx->idx_dest.ensure(std::max<size_t>(len, 1)));
    if (ctx->g.relabeled) { // per-vertex info permuted to internal order, destinations mapped to internal ids
        DevBuf<u64> t64;
        DevBuf<int32_t> t32;
        CK(t64.ensure(n));
        CK(t32.ensure(std::max<size_t>(len, 1)));
        CK(cudaMemcpyAsync(t64.p, offsets, sizeof(u64) * n, cudaMemcpyHostToDevice, ctx->stream));
        gather_kernel<u64><<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(n, t64.p, ctx->g.new2old, ctx->idx_off.p);
        CKL();
        CK(cudaMemcpyAsync(t64.p, counts, sizeof(u64) * n, cudaMemcpyHostToDevice, ctx->stream));
        gather_kernel<u64><<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(n, t64.p, ctx->g.new2old, ctx->idx_cnt.p);
        CKL();
        CK(cudaMemcpyAsync(t32.p, dest, sizeof(int32_t) * len, cudaMemcpyHostToDevice, ctx->stream));
        if (len) {
            map_ids_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(len, t32.p, ctx->g.old2new, ctx->idx_dest.p);
            CKL();
        }
        CK(cudaStreamSynchronize(ctx->stream));
        t64.release(); t32.release();
    } else {
        CK(cudaMemcpyAsync(ctx->idx_off.p, offsets, sizeof(u64) * n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->idx_cnt.p, counts, sizeof(u64) * n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->idx_dest.p, dest, sizeof(int32_t) * len, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->has_index = true;
    return FORA_OK;
}

// =============================================================================================
// top-k of a dense vector and power iteration live in topk.cuh
// =============================================================================================
extern "C" int fora_topk_of(fora_ctx* ctx, const double* ppr, uint32_t k, int32_t* nodes, double* values) {
    if (!ctx || !ctx->g.n) return ctx ? ctx->fail(FORA_EINVAL, "no graph") : FORA_EINVAL;
    if (!ppr || !nodes || !values || k == 0) return ctx->fail(FORA_EINVAL, "bad arguments");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->g.n;
    CK(ctx->scratchd.ensure(n));
    CK(cudaMemcpyAsync(ctx->scratchd.p, ppr, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    const int zero = 0;
    SelResult r;
    int rc = select_batch(ctx, ctx->scratchd.p, 0, &zero, 1, k, true, &r);
    if (rc) return rc;
    if (r.out_count) {
        CK(cudaMemcpyAsync(nodes, ctx->sel_on.p, sizeof(int32_t) * r.out_count, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(values, ctx->sel_ov.p, sizeof(double) * r.out_count, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    for (uint32_t j = r.out_count; j < k; ++j) { nodes[j] = 0; values[j] = 0.0; }
    return FORA_OK;
}

extern "C" int fora_power_iteration(fora_ctx* ctx, int32_t source, int iters, double* ppr) {
    if (!ctx || !ctx->g.n) return ctx ? ctx->fail(FORA_EINVAL, "no graph") : FORA_EINVAL;
    if (source < 0 || source >= ctx->g.n || iters < 0 || !ppr) return ctx->fail(FORA_EINVAL, "bad arguments");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->g.n;
    CK(ctx->scratchd.ensure(3 * n + 2));
    cudaError_t e;
    if (ctx->g.off32) e = power_iteration_device<u32>(ctx->stream, ctx->num_sms, CsrView<u32>{ctx->g.out_ptr32, ctx->g.out_col}, ctx->g.n, to_internal(ctx, source), iters, ctx->p.alpha, ctx->scratchd.p, &ctx->launches);
    else e = power_iteration_device<int64_t>(ctx->stream, ctx->num_sms, CsrView<int64_t>{ctx->g.out_ptr64, ctx->g.out_col}, ctx->g.n, to_internal(ctx, source), iters, ctx->p.alpha, ctx->scratchd.p, &ctx->launches);
    if (e != cudaSuccess) return ctx->fail(FORA_ECUDA, std::string("power iteration: ") + cudaGetErrorString(e));
    CK(ctx->stage.ensure(n));
    int prc = vec_to_original(ctx, ctx->scratchd.p, ctx->stage.p, n);
    if (prc) return prc;
    CK(cudaMemcpyAsync(ppr, ctx->stage.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FORA_OK;
}

// =============================================================================================
// not yet wired (fail loudly rather than fall back)
// =============================================================================================

// development aid: copy the per-level trace of the last push launch (FORA_PUSH_TRACE=1); returns levels
extern "C" int fora_debug_push_trace(fora_ctx* ctx, uint64_t* out, int cap_levels) {
    if (!ctx || !ctx->trace_on) return 0;
    PushCtl h;
    cudaMemcpy(&h, ctx->ctl.p, sizeof h, cudaMemcpyDeviceToHost);
    const int lv = std::min<int>({(int)h.levels_run, cap_levels, 4096});
    cudaMemcpy(out, ctx->trace.p, sizeof(u64) * 4 * std::min(cap_levels, 4096), cudaMemcpyDeviceToHost); // (push3 keeps more behind the levels)
    return lv;
}

// =============================================================================================
// several GPUs on one query (NCCL)
// =============================================================================================
#include "group.cuh"
