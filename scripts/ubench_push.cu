// Microbenchmarks behind the round-2 push design (development aid, not product code):
//   foot   random fp64 ATOM/RED throughput vs footprint around the L2 capacity (1 .. 5 LJ-size slot vectors)
//   bar    latency of a grid-wide barrier: cooperative_groups grid.sync vs a hand-rolled counter/generation barrier
//   chain  latency of dependent L2 accesses (pointer chase, atomic-with-return chain)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -rdc=false -o scripts/ubench_push scripts/ubench_push.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
namespace cg = cooperative_groups;
typedef unsigned long long u64;
typedef unsigned int u32;
__device__ __forceinline__ u32 hash32(u32 x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// ---- foot -------------------------------------------------------------------------------------
template <int MODE, int INFLIGHT>
__global__ void foot_kernel(double* a, u32 n, u64 total, double* sink) {
    const u64 tid = blockIdx.x * (u64)blockDim.x + threadIdx.x, gs = gridDim.x * (u64)blockDim.x;
    double acc = 0;
    for (u64 i = tid; i < total; i += gs * INFLIGHT) {
        u32 j[INFLIGHT];
#pragma unroll
        for (int k = 0; k < INFLIGHT; ++k) j[k] = (u32)(((u64)hash32((u32)(i + k * gs) * 2654435761u + 12345u) * n) >> 32);
#pragma unroll
        for (int k = 0; k < INFLIGHT; ++k) {
            if (i + k * gs < total) {
                if (MODE == 0) acc += atomicAdd(&a[j[k]], 1e-9);
                else atomicAdd(&a[j[k]], 1e-9);
            }
        }
    }
    if (acc == 123.456) *sink = acc;
}
template <int MODE, int INFLIGHT>
static void foot(const char* name, double* a, u32 n, u64 total, double* sink) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    foot_kernel<MODE, INFLIGHT><<<148 * 2, 1024>>>(a, n, total / 4, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    foot_kernel<MODE, INFLIGHT><<<148 * 2, 1024>>>(a, n, total, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-22s inflight=%d n=%9u (%6.1f MB): %8.2f G ops/s\n", name, INFLIGHT, n, n * 8.0 / 1e6, total / ms / 1e6);
}

// ---- bar --------------------------------------------------------------------------------------
__global__ void bar_cg_kernel(int iters, u64* out) {
    cg::grid_group grid = cg::this_grid();
    u64 t0 = 0, t1 = 0;
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int i = 0; i < iters; ++i) grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        out[0] = t1 - t0;
    }
}
// counter + generation: the last arriver resets the counter and bumps the generation, everybody else spins on it
__device__ __forceinline__ void hand_barrier(u32* count, volatile u32* gen, u32 nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const u32 g = *gen;
        __threadfence();
        if (atomicAdd(count, 1u) == nblocks - 1) {
            *count = 0;
            __threadfence();
            atomicAdd((u32*)gen, 1u);
        } else {
            while (*gen == g) {}
        }
        __threadfence();
    }
    __syncthreads();
}
__global__ void bar_hand_kernel(int iters, u64* out, u32* count, u32* gen) {
    u64 t0 = 0, t1 = 0;
    hand_barrier(count, gen, gridDim.x);
    if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int i = 0; i < iters; ++i) hand_barrier(count, gen, gridDim.x);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        out[0] = t1 - t0;
    }
}

// ---- chain ------------------------------------------------------------------------------------
__global__ void chase_kernel(const u32* next, int iters, u64* out, u32* sink) {
    u32 p = threadIdx.x * 977u;
    u64 t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int i = 0; i < iters; ++i) p = __ldcg(&next[p]);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (p == 0xffffffffu) *sink = p;
}
__global__ void atom_chain_kernel(double* a, u32 n, int iters, u64* out, double* sink) {
    u32 p = threadIdx.x * 977u + 13u;
    double acc = 0;
    u64 t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int i = 0; i < iters; ++i) {
        const double o = atomicAdd(&a[p % n], 1e-9);
        acc += o;
        p = hash32(p + (u32)__double_as_longlong(o));
    }
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (acc == 123.456) *sink = acc;
}
__global__ void fill_next(u32* next, u32 n) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) next[i] = (u32)(((u64)hash32(i * 2654435761u + 7u) * n) >> 32);
}

int main(int argc, char** argv) {
    const char* what = argc > 1 ? argv[1] : "all";
    double *a, *sink;
    u64* out;
    u32 *count, *gen, *next;
    const u32 nmax = 4847571u * 6;
    cudaMalloc(&a, nmax * 8ull); cudaMalloc(&sink, 8); cudaMalloc(&out, 64); cudaMalloc(&count, 4); cudaMalloc(&gen, 4);
    cudaMemset(a, 0, nmax * 8ull); cudaMemset(count, 0, 4); cudaMemset(gen, 0, 4);
    if (!strcmp(what, "all") || !strcmp(what, "foot")) {
        const u64 total = 1ull << 29;
        for (u32 q : {4u, 6u, 8u, 10u, 12u, 14u, 16u, 20u, 24u}) { // quarter slots: 1, 1.5, 2, 2.5, 3, 3.5, 4, 5, 6 slot vectors
            const u32 n = (u32)(4847571ull * q / 4);
            foot<0, 1>("ATOM.f64 (return)", a, n, total, sink);
            foot<0, 4>("ATOM.f64 (return)", a, n, total, sink);
            foot<1, 1>("RED.f64", a, n, total, sink);
        }
    }
    if (!strcmp(what, "all") || !strcmp(what, "bar")) {
        const int iters = 2000;
        u64 h;
        for (int cfg = 0; cfg < 3; ++cfg) {
            const int grid = cfg == 0 ? 148 : (cfg == 1 ? 296 : 592), block = cfg == 0 ? 1024 : (cfg == 1 ? 512 : 256);
            void* args[] = {(void*)&iters, (void*)&out};
            cudaError_t e = cudaLaunchCooperativeKernel((void*)bar_cg_kernel, dim3(grid), dim3(block), args, 0, 0);
            cudaDeviceSynchronize();
            cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            printf("grid.sync          %4d x %4d: %7.3f us per barrier (%s)\n", grid, block, h / 1e3 / iters, cudaGetErrorString(e));
            void* args2[] = {(void*)&iters, (void*)&out, (void*)&count, (void*)&gen};
            e = cudaLaunchCooperativeKernel((void*)bar_hand_kernel, dim3(grid), dim3(block), args2, 0, 0);
            cudaDeviceSynchronize();
            cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            printf("hand-rolled        %4d x %4d: %7.3f us per barrier (%s)\n", grid, block, h / 1e3 / iters, cudaGetErrorString(e));
        }
    }
    if (!strcmp(what, "all") || !strcmp(what, "chain")) {
        u64 h;
        for (u32 mb : {16u, 64u, 512u}) {
            const u32 n = mb * 262144u;
            cudaMalloc(&next, n * 4ull);
            fill_next<<<592, 256>>>(next, n);
            chase_kernel<<<1, 32>>>(next, 2000, out, (u32*)sink);
            chase_kernel<<<1, 32>>>(next, 20000, out, (u32*)sink);
            cudaDeviceSynchronize();
            cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            printf("pointer chase (ld.cg) over %4u MB: %7.1f ns per dependent load\n", mb, h / 20000.0);
            cudaFree(next);
        }
        for (u32 n : {4847571u, 4847571u * 6}) {
            atom_chain_kernel<<<1, 32>>>(a, n, 2000, out, sink);
            atom_chain_kernel<<<1, 32>>>(a, n, 20000, out, sink);
            cudaDeviceSynchronize();
            cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            printf("ATOM.f64 return chain over %5.1f MB: %7.1f ns per dependent atomic\n", n * 8.0 / 1e6, h / 20000.0);
        }
    }
    return 0;
}
