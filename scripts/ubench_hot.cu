// Microbenchmark behind the round-2 push scatter (development aid, not product code).
// Question: what bounds fp64 ATOM-with-return scatters whose destinations follow the graph's in-degree distribution
// (shifted power law, hot vertices contiguous after the engine's relabelling) -- the L2 atomic rate, the DRAM sector
// rate, or serialisation on the few hottest 128-byte lines?
//   dist    one 39 MB vector (L2-resident): uniform vs power-law destinations, with / without a shared-memory
//           accumulator for the first K (hottest) indices that is flushed once per tile
//   sweep   S slot vectors swept slot-major by tiles taken round-robin (the push kernel's phase-B schedule): power-law
//           destinations, with / without an L2 bulk prefetch of the slot the sweep reaches next, with / without the accumulator
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_hot scripts/ubench_hot.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
typedef unsigned long long u64;
typedef unsigned int u32;
__device__ __forceinline__ u32 hash32(u32 x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// destinations: mode 0 uniform, mode 1 rank ~ (rank + 50)^(-1/1.3) (host_util.cpp synth_edges), rank == index
__global__ void gen_kernel(u32* dst, u64 total, u32 n, int mode, u32 salt) {
    const double beta = 1.0 / 1.3, shift = 50.0, e1 = 1.0 - beta;
    const double a = pow(shift, e1), b = pow((double)n + shift, e1);
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < total; i += gridDim.x * (u64)blockDim.x) {
        const u32 h1 = hash32((u32)i * 2654435761u + salt), h2 = hash32(h1 ^ (u32)(i >> 32) ^ 0x9e3779b9u);
        const double u = ((double)h1 * 4294967296.0 + (double)h2) * (1.0 / 18446744073709551616.0);
        u32 r;
        if (mode == 0) r = (u32)(u * n);
        else {
            double x = pow(u * (b - a) + a, 1.0 / e1) - shift;
            r = x < 0 ? 0u : (u32)x;
        }
        dst[i] = r < n ? r : n - 1;
    }
}

__device__ __forceinline__ void prefetch_l2_bulk(const void* p, u32 bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// One tile = `tile` consecutive edges of the slot-major edge line; CTA b takes tiles b, b+grid, ...
// AGG: indices < K accumulate in shared memory and are flushed with one ATOM each at the end of the tile.
template <int K, int INFLIGHT>
__global__ void __launch_bounds__(512, 2) sweep_kernel(double* res, const u32* __restrict__ dst, u32 n, u64 per_slot, int slots, u32 tile,
                                                       int prefetch, double* sink) {
    extern __shared__ double hot[];
    const u64 E = per_slot * slots;
    const u64 ntiles = (E + tile - 1) / tile;
    const u64 tiles_per_slot = (per_slot + tile - 1) / tile;
    double acc = 0;
    for (u64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const u64 lo = t * tile, hi = min(E, lo + tile);
        const int slot = (int)(lo / per_slot);
        double* r = res + (size_t)slot * n;
        if (K) {
            for (int k = threadIdx.x; k < K; k += blockDim.x) hot[k] = 0.0;
            __syncthreads();
        }
        if (prefetch && threadIdx.x == 0) {
            // this tile's share of the vector the sweep reaches `prefetch` slots later
            const int ps = slot + prefetch;
            if (ps < slots) {
                const u64 j = t % tiles_per_slot;
                const size_t bytes = (size_t)n * 8;
                size_t b0 = (bytes * j / tiles_per_slot) & ~(size_t)127, b1 = (bytes * (j + 1) / tiles_per_slot) & ~(size_t)127;
                const char* base = (const char*)(res + (size_t)ps * n);
                for (size_t o = b0; o < b1; o += 32768) prefetch_l2_bulk(base + o, (u32)min((size_t)32768, b1 - o));
            }
        }
        for (u64 x0 = lo + threadIdx.x; x0 < hi; x0 += (u64)blockDim.x * INFLIGHT) {
            u32 j[INFLIGHT];
#pragma unroll
            for (int k = 0; k < INFLIGHT; ++k) {
                const u64 x = x0 + (u64)k * blockDim.x;
                j[k] = x < hi ? __ldcs(&dst[x - (u64)slot * per_slot + (u64)((slot * 1237) & 4095)]) : 0xffffffffu;
            }
#pragma unroll
            for (int k = 0; k < INFLIGHT; ++k) {
                if (j[k] == 0xffffffffu) continue;
                if (K && j[k] < (u32)K) atomicAdd(&hot[j[k]], 1e-9);
                else acc += atomicAdd(&r[j[k]], 1e-9);
            }
        }
        if (K) {
            __syncthreads();
            for (int k = threadIdx.x; k < K; k += blockDim.x) {
                const double v = hot[k];
                if (v != 0.0) acc += atomicAdd(&r[k], v);
            }
            __syncthreads();
        }
    }
    if (acc == 123.456) *sink = acc;
}

template <int K, int INFLIGHT>
static double run(const char* name, double* res, const u32* dst, u32 n, u64 per_slot, int slots, u32 tile, int prefetch, double* sink,
                  int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const size_t smem = K * sizeof(double);
    cudaFuncSetAttribute(sweep_kernel<K, INFLIGHT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sweep_kernel<K, INFLIGHT><<<148 * 2, 512, smem>>>(res, dst, n, per_slot, slots, tile, prefetch, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) sweep_kernel<K, INFLIGHT><<<148 * 2, 512, smem>>>(res, dst, n, per_slot, slots, tile, prefetch, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const cudaError_t err = cudaGetLastError();
    const double rate = (double)per_slot * slots * reps / ms / 1e6;
    printf("%-8s K=%5d inflight=%d slots=%2d edges/slot=%8llu tile=%6u prefetch=%d: %8.2f G atomics/s (%s)\n", name, K, INFLIGHT, slots,
           per_slot, tile, prefetch, rate, cudaGetErrorString(err));
    fflush(stdout);
    return rate;
}

int main(int argc, char** argv) {
    const char* what = argc > 1 ? argv[1] : "all";
    const u32 n = 4847571;
    const int S = 48;
    const u64 maxper = 4u << 20;
    double *res, *sink;
    u32 *du, *dp;
    cudaMalloc(&res, (size_t)S * n * 8);
    cudaMemset(res, 0, (size_t)S * n * 8);
    cudaMalloc(&sink, 8);
    cudaMalloc(&du, (maxper + 8192) * 4);
    cudaMalloc(&dp, (maxper + 8192) * 4);
    gen_kernel<<<1024, 256>>>(du, maxper + 8192, n, 0, 11u);
    gen_kernel<<<1024, 256>>>(dp, maxper + 8192, n, 1, 23u);
    cudaDeviceSynchronize();
    {   // how hot is hot: share of the first 16 / 1024 / 16384 indices in the power-law stream
        u32* h = (u32*)malloc(maxper * 4);
        cudaMemcpy(h, dp, maxper * 4, cudaMemcpyDeviceToHost);
        u64 c16 = 0, c1k = 0, c16k = 0, c0 = 0;
        for (u64 i = 0; i < maxper; ++i) { c0 += h[i] == 0; c16 += h[i] < 16; c1k += h[i] < 1024; c16k += h[i] < 16384; }
        printf("power-law stream: index 0 %.4f %%, first 16 %.3f %%, first 1024 %.2f %%, first 16384 %.2f %%\n", 100.0 * c0 / maxper,
               100.0 * c16 / maxper, 100.0 * c1k / maxper, 100.0 * c16k / maxper);
        free(h);
    }
    if (!strcmp(what, "dist") || !strcmp(what, "all")) {
        // one L2-resident vector, 4 M edges per pass, 16 passes
        run<0, 2>("uniform", res, du, n, maxper, 1, 16384, 0, sink, 16);
        run<0, 2>("powerlaw", res, dp, n, maxper, 1, 16384, 0, sink, 16);
        run<0, 4>("powerlaw", res, dp, n, maxper, 1, 16384, 0, sink, 16);
        run<256, 2>("powerlaw", res, dp, n, maxper, 1, 16384, 0, sink, 16);
        run<2048, 2>("powerlaw", res, dp, n, maxper, 1, 16384, 0, sink, 16);
        run<8192, 2>("powerlaw", res, dp, n, maxper, 1, 16384, 0, sink, 16);
        run<2048, 2>("powerlaw", res, dp, n, maxper, 1, 32768, 0, sink, 16);
        run<8192, 2>("powerlaw", res, dp, n, maxper, 1, 32768, 0, sink, 16);
        run<8192, 4>("powerlaw", res, dp, n, maxper, 1, 32768, 0, sink, 16);
        run<2048, 2>("uniform", res, du, n, maxper, 1, 16384, 0, sink, 16);
    }
    if (!strcmp(what, "sweep") || !strcmp(what, "all")) {
        for (u64 per : {4ull << 20, 1ull << 20, 256ull << 10}) {
            for (u32 tile : {8192u, 32768u}) {
                run<0, 2>("uniform", res, du, n, per, S, tile, 0, sink, 2);
                run<0, 2>("powerlaw", res, dp, n, per, S, tile, 0, sink, 2);
                run<0, 2>("powerlaw", res, dp, n, per, S, tile, 1, sink, 2);
                run<2048, 2>("powerlaw", res, dp, n, per, S, tile, 0, sink, 2);
                run<2048, 2>("powerlaw", res, dp, n, per, S, tile, 1, sink, 2);
                run<2048, 2>("powerlaw", res, dp, n, per, S, tile, 2, sink, 2);
                run<8192, 2>("powerlaw", res, dp, n, per, S, tile, 1, sink, 2);
            }
        }
    }
    return 0;
}
