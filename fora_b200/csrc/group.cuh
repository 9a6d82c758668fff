// fora_b200/csrc/group.cuh -- several GPUs of one box working on ONE whole-graph SSPPR query (BASELINE config 5).
//
// The reference computes a query in one function on one core (fora_query_basic, /root/reference/query.h:841-907: push
// algo.h:954-1093, walks query.h:334-413).  Here: GPU 0 pushes; the push state the other GPUs need -- the compacted list of
// (vertex, residue) pairs, not the two dense vectors -- is broadcast with ncclBroadcast; every GPU rebuilds the SAME walk plan
// from it and walks chunk range g of G (Philox is keyed by source and walk index, so the split does not change a single
// destination); the dense fp64 vectors are summed with one ncclAllReduce over NVLink / NVSwitch.  No host round trip of the
// push state.  One host thread drives all GPUs (NCCL group calls); NCCL is loaded at run time (dlopen), so the library itself
// does not depend on it.  Included at the end of engine.cu: uses the engine's internals.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

struct NcclApi {
    void* so = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err) {
        if (so) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            so = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (so) break;
        }
        if (!so) { err = std::string("NCCL not found: ") + dlerror(); return false; }
#define FORA_NCCL_SYM(field, sym) field = (decltype(field))dlsym(so, sym); if (!field) { err = std::string("NCCL symbol missing: ") + sym; return false; }
        FORA_NCCL_SYM(CommInitAll, "ncclCommInitAll");
        FORA_NCCL_SYM(CommDestroy, "ncclCommDestroy");
        FORA_NCCL_SYM(Broadcast, "ncclBroadcast");
        FORA_NCCL_SYM(AllReduce, "ncclAllReduce");
        FORA_NCCL_SYM(GroupStart, "ncclGroupStart");
        FORA_NCCL_SYM(GroupEnd, "ncclGroupEnd");
        FORA_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef FORA_NCCL_SYM
        return true;
    }
};
static NcclApi g_nccl;

struct fora_group {
    std::vector<fora_ctx*> ctx;
    std::vector<int> dev;
    std::vector<ncclComm_t> comm;
    std::vector<DevBuf<int32_t> > ids;   // compacted residue list, per GPU
    std::vector<DevBuf<double> > vals;
    std::vector<cudaEvent_t> ev;         // 4 per GPU
    DevBuf<u32> cnt0;                    // on GPU 0
    std::string err;
    int fail(int code, const std::string& m) { err = m; return code; }
};
static std::string g_group_error;

#define GCK(call)                                                                                               \
    do {                                                                                                        \
        cudaError_t e__ = (call);                                                                               \
        if (e__ != cudaSuccess) return grp->fail(FORA_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)
#define GNK(call)                                                                                                    \
    do {                                                                                                             \
        ncclResult_t r__ = (call);                                                                                   \
        if (r__ != ncclSuccess) return grp->fail(FORA_ECUDA, std::string(#call) + ": " + g_nccl.GetErrorString(r__)); \
    } while (0)

extern "C" int fora_group_create(int n_gpus, const int* devices, uint64_t seed, fora_group** out) {
    if (!out || n_gpus < 1) return FORA_EINVAL;
    *out = nullptr;
    fora_group* grp = new fora_group();
    for (int i = 0; i < n_gpus; ++i) {
        const int d = devices ? devices[i] : i;
        fora_ctx* c = nullptr;
        int rc = fora_ctx_create(d, seed, &c); // the same seed everywhere: the GPUs share one Philox key space
        if (rc) {
            g_group_error = fora_last_error(nullptr);
            for (auto* x : grp->ctx) fora_ctx_destroy(x);
            delete grp;
            return rc;
        }
        fora_ctx_set_slots(c, 1);
        grp->ctx.push_back(c);
        grp->dev.push_back(d);
    }
    grp->ids.resize(n_gpus);
    grp->vals.resize(n_gpus);
    if (n_gpus > 1) {
        if (!g_nccl.load(g_group_error)) {
            for (auto* x : grp->ctx) fora_ctx_destroy(x);
            delete grp;
            return FORA_ECUDA;
        }
        grp->comm.resize(n_gpus);
        ncclResult_t r = g_nccl.CommInitAll(grp->comm.data(), n_gpus, grp->dev.data());
        if (r != ncclSuccess) {
            g_group_error = std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r);
            for (auto* x : grp->ctx) fora_ctx_destroy(x);
            delete grp;
            return FORA_ECUDA;
        }
    }
    grp->ev.resize(4 * (size_t)n_gpus);
    for (int i = 0; i < n_gpus; ++i) {
        cudaSetDevice(grp->dev[i]);
        for (int j = 0; j < 4; ++j) cudaEventCreate(&grp->ev[4 * i + j]);
    }
    *out = grp;
    return FORA_OK;
}
extern "C" void fora_group_destroy(fora_group* grp) {
    if (!grp) return;
    for (size_t i = 0; i < grp->ctx.size(); ++i) {
        cudaSetDevice(grp->dev[i]);
        cudaDeviceSynchronize();
        grp->ids[i].release();
        grp->vals[i].release();
        if (i == 0) grp->cnt0.release();
        for (int j = 0; j < 4; ++j) cudaEventDestroy(grp->ev[4 * i + j]);
    }
    for (auto& c : grp->comm) g_nccl.CommDestroy(c);
    for (auto* x : grp->ctx) fora_ctx_destroy(x);
    delete grp;
}
extern "C" int fora_group_size(fora_group* grp) { return grp ? (int)grp->ctx.size() : FORA_EINVAL; }
extern "C" fora_ctx* fora_group_ctx(fora_group* grp, int i) { return (grp && i >= 0 && i < (int)grp->ctx.size()) ? grp->ctx[(size_t)i] : nullptr; }
extern "C" const char* fora_group_last_error(fora_group* grp) { return grp ? grp->err.c_str() : g_group_error.c_str(); }

// scatter the compacted list back into a zeroed dense vector
__global__ void group_scatter_kernel(u32 cnt, const int32_t* __restrict__ ids, const double* __restrict__ vals, double* __restrict__ dense) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) dense[ids[i]] = vals[i];
}

extern "C" int fora_group_query_split(fora_group* grp, int32_t source, uint32_t query_id, double* ppr, fora_query_stat* stat, fora_split_timing* tm) {
    if (!grp) return FORA_EINVAL;
    const int G = (int)grp->ctx.size();
    fora_ctx* c0 = grp->ctx[0];
    if (!c0->g.n) return grp->fail(FORA_EINVAL, "no graph uploaded");
    const size_t n = (size_t)c0->g.n;
    if (source < 0 || source >= c0->g.n) return grp->fail(FORA_EINVAL, "source out of range");
    for (int i = 0; i < G; ++i) {
        fora_ctx* c = grp->ctx[(size_t)i];
        if (c->g.n != c0->g.n || c->g.n_edges != c0->g.n_edges) return grp->fail(FORA_EINVAL, "every GPU of the group must hold the same graph");
        if (c->slots != 1) fora_ctx_set_slots(c, 1);
        int rc = require_ready(c, c->p.omega);
        if (rc) return grp->fail(rc, c->err);
    }
    auto cfail = [&](fora_ctx* c, int rc) { return grp->fail(rc, c->err); };
    int rc;
    // ---- GPU 0: push.  G GPUs make a walk G times cheaper, so --balanced stops the push earlier (query.h:826-839 with the walk cost / G)
    GCK(cudaSetDevice(grp->dev[0]));
    GCK(cudaEventRecord(grp->ev[0], c0->stream));
    const fora_params keep = c0->p;
    if (c0->p.balanced && G > 1) c0->p.cost_walk = keep.cost_walk / G;
    c0->h_meta->source[0] = to_internal(c0, source);
    c0->h_meta->qid[0] = query_id;
    double fr = 0;
    u64 rounds = 0;
    rc = push_wave(c0, 1, nullptr, &fr, &rounds);
    c0->p = keep;
    if (rc) return cfail(c0, rc);
    if ((rc = meta_d2h_sync(c0))) return cfail(c0, rc);
    const int state0 = c0->h_meta->state[0];
    const double rsum = state0 == 1 ? c0->h_meta->rsum[0] : 0.0;
    fora_query_stat st0;
    fill_stat(c0, 0, fr, rounds, &st0);
    GCK(cudaEventRecord(grp->ev[1], c0->stream));
    // ---- compact (vertex, residue) on GPU 0, broadcast, rebuild the dense residue on the other GPUs
    u32 cnt = 0;
    if (G > 1 && rsum > 0.0) {
        GCK(grp->cnt0.ensure(MAX_SLOTS));
        GCK(grp->ids[0].ensure(n));
        GCK(grp->vals[0].ensure(n));
        GCK(cudaMemsetAsync(grp->cnt0.p, 0, sizeof(u32) * MAX_SLOTS, c0->stream));
        const int gx = std::max(1, std::min(c0->num_sms * 8, (int)((n + 2047) / 2048)));
        sparse_out_kernel<<<dim3(gx, 1), 256, 0, c0->stream>>>(c0->g.n, c0->residue.p, nullptr, 4.9406564584124654e-324, (u32)std::min<size_t>(n, 0xfffffff0u), 0,
                                                             grp->ids[0].p, grp->vals[0].p, grp->cnt0.p); // every residue > 0, internal ids
        GCK(cudaMemcpyAsync(&cnt, grp->cnt0.p, sizeof(u32), cudaMemcpyDeviceToHost, c0->stream));
        GCK(cudaStreamSynchronize(c0->stream));
        for (int i = 1; i < G; ++i) {
            GCK(cudaSetDevice(grp->dev[(size_t)i]));
            GCK(grp->ids[(size_t)i].ensure(n));
            GCK(grp->vals[(size_t)i].ensure(n));
        }
        GNK(g_nccl.GroupStart());
        for (int i = 0; i < G; ++i) {
            GNK(g_nccl.Broadcast(grp->ids[0].p, grp->ids[(size_t)i].p, cnt, ncclInt32, 0, grp->comm[(size_t)i], grp->ctx[(size_t)i]->stream));
            GNK(g_nccl.Broadcast(grp->vals[0].p, grp->vals[(size_t)i].p, cnt, ncclFloat64, 0, grp->comm[(size_t)i], grp->ctx[(size_t)i]->stream));
        }
        GNK(g_nccl.GroupEnd());
        for (int i = 1; i < G; ++i) {
            fora_ctx* c = grp->ctx[(size_t)i];
            GCK(cudaSetDevice(grp->dev[(size_t)i]));
            GCK(cudaMemsetAsync(c->residue.p, 0, sizeof(double) * n, c->stream));
            GCK(cudaMemsetAsync(c->reserve.p, 0, sizeof(double) * n, c->stream)); // parts > 0 start from zero: the sum over GPUs is the PPR vector
            if (cnt) group_scatter_kernel<<<std::max(1, std::min(c->num_sms * 8, (int)((cnt + 255) / 256))), 256, 0, c->stream>>>(cnt, grp->ids[(size_t)i].p, grp->vals[(size_t)i].p, c->residue.p);
        }
    }
    GCK(cudaSetDevice(grp->dev[0]));
    GCK(cudaEventRecord(grp->ev[2], c0->stream));
    // ---- every GPU walks its chunk range of the same plan
    const int parts = rsum > 0.0 ? G : 1;
    for (int i = 0; i < parts; ++i) {
        fora_ctx* c = grp->ctx[(size_t)i];
        GCK(cudaSetDevice(grp->dev[(size_t)i]));
        SlotMeta* h = c->h_meta;
        if (i > 0) {
            memset(h, 0, sizeof *h);
            for (int s = 0; s < MAX_SLOTS; ++s) h->source[s] = -1;
            h->source[0] = 0;
        }
        h->qid[0] = query_id;
        h->state[0] = rsum == 0.0 ? 2 : 1;
        h->rsum[0] = rsum;
        h->nwalk[0] = h->hops[0] = h->idx_hits[0] = h->nsrc[0] = 0;
        if ((rc = meta_h2d(c))) return cfail(c, rc);
        GCK(cudaEventRecord(grp->ev[4 * (size_t)i + 3], c->stream)); // (re-used below: start of this GPU's walk phase)
        if ((rc = walk_wave(c, c->reserve.p, 0, c->p.opt, c->p.opt, 0, nullptr, (u32)i, (u32)parts))) return cfail(c, rc);
    }
    std::vector<cudaEvent_t> walk_end((size_t)parts);
    float walk_ms = 0;
    for (int i = 0; i < parts; ++i) {
        fora_ctx* c = grp->ctx[(size_t)i];
        GCK(cudaSetDevice(grp->dev[(size_t)i]));
        GCK(cudaEventCreate(&walk_end[(size_t)i]));
        GCK(cudaEventRecord(walk_end[(size_t)i], c->stream));
    }
    // ---- sum the dense vectors
    if (parts > 1) {
        GNK(g_nccl.GroupStart());
        for (int i = 0; i < G; ++i) GNK(g_nccl.AllReduce(grp->ctx[(size_t)i]->reserve.p, grp->ctx[(size_t)i]->reserve.p, n, ncclFloat64, ncclSum, grp->comm[(size_t)i], grp->ctx[(size_t)i]->stream));
        GNK(g_nccl.GroupEnd());
    }
    GCK(cudaSetDevice(grp->dev[0]));
    cudaEvent_t red_end;
    GCK(cudaEventCreate(&red_end));
    GCK(cudaEventRecord(red_end, c0->stream));
    if (ppr) {
        GCK(c0->scratchd.ensure(n));
        if ((rc = vec_to_original(c0, c0->reserve.p, c0->scratchd.p, n))) return cfail(c0, rc);
        GCK(cudaMemcpyAsync(ppr, c0->scratchd.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c0->stream));
    }
    cudaEvent_t all_end;
    GCK(cudaEventCreate(&all_end));
    GCK(cudaEventRecord(all_end, c0->stream));
    u64 walks = 0, hops = 0, nsrc = 0;
    for (int i = 0; i < parts; ++i) {
        fora_ctx* c = grp->ctx[(size_t)i];
        GCK(cudaSetDevice(grp->dev[(size_t)i]));
        if ((rc = meta_d2h_sync(c))) return cfail(c, rc);
        hops += c->h_meta->hops[0];
        walks = c->h_meta->nwalk[0]; // the plan (and so its walk count) is the same on every GPU; each walked 1/G of it
        nsrc = c->h_meta->nsrc[0];
        float ms = 0;
        cudaEventElapsedTime(&ms, grp->ev[4 * (size_t)i + 3], walk_end[(size_t)i]);
        walk_ms = std::max(walk_ms, ms);
        cudaEventDestroy(walk_end[(size_t)i]);
    }
    GCK(cudaSetDevice(grp->dev[0]));
    GCK(cudaEventSynchronize(all_end));
    if (stat) {
        *stat = st0;
        stat->n_walks = walks;
        stat->walk_hops = hops;
        stat->n_sources = nsrc;
    }
    if (tm) {
        memset(tm, 0, sizeof *tm);
        cudaEventElapsedTime(&tm->total_ms, grp->ev[0], all_end);
        cudaEventElapsedTime(&tm->push_ms, grp->ev[0], grp->ev[1]);
        cudaEventElapsedTime(&tm->bcast_ms, grp->ev[1], grp->ev[2]);
        tm->walk_ms = walk_ms; // plan + walk kernels, slowest GPU
        float t_walk_red = 0;
        cudaEventElapsedTime(&t_walk_red, grp->ev[2], red_end);
        tm->reduce_ms = std::max(0.0f, t_walk_red - walk_ms);
        tm->bcast_bytes = (uint64_t)cnt * 12u;
        tm->reduce_bytes = parts > 1 ? (uint64_t)n * 8u : 0u;
        tm->n_gpus = (uint32_t)G;
    }
    cudaEventDestroy(red_end);
    cudaEventDestroy(all_end);
    return FORA_OK;
}
