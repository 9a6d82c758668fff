// Third sweep microbenchmark (development aid): the grid moves through the S slot vectors in LOCKSTEP (a grid barrier after every
// group of K vectors), so that at most K (+K) vectors are live in the L2 at any time.
//   B-only : E atomics with return on every vector of the group
//   A|B    : software pipeline -- in the same barrier interval the grid does the F exchanges (phase A) of the NEXT group and the E
//            adds (phase B) of the current one
//   RED|scan: in one interval the grid fires the E adds of vector s as RED (no return value) and scans vector s-1 densely
//            (every element read once, elements above a threshold zeroed and counted: the next frontier found without return values)
// usage: ubench_sweep3 S E K mode(0 B-only, 1 A|B, 2 RED|scan, 3 RED only) [blocks_per_sm] [threads]
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;
typedef unsigned long long u64;
typedef unsigned int u32;
__device__ __forceinline__ u32 hash32(u32 x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ u32 target(u64 op, u32 n) {
    const u32 h = hash32((u32)op * 2654435761u + (u32)(op >> 32) * 40503u + 12345u);
    return (u32)(((u64)h * n) >> 32);
}
__global__ void sweep(double* a, u32 n, int S, u64 E, u64 F, int K, int mode, u32 salt, double* sink) {
    cg::grid_group grid = cg::this_grid();
    double acc = 0;
    const u64 gs = (u64)gridDim.x * blockDim.x, tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (mode >= 2) {
        u32 found = 0;
        for (int s = 0; s <= S; ++s) {
            if (s > 0 && mode == 2) { // dense scan of vector s-1: contiguous share per CTA, 16-byte loads
                double* v = a + (size_t)(s - 1) * n; // (slot vectors are only 8-byte aligned: n is odd)
                const u32 per = (n + gridDim.x - 1) / gridDim.x, b = blockIdx.x * per, e = min(n, b + per);
                for (u32 i = b + threadIdx.x; i < e; i += 2 * blockDim.x) {
                    const double r0 = v[i];
                    const double r1 = i + blockDim.x < e ? v[i + blockDim.x] : 0.0;
                    if (r0 >= 3e-9) { v[i] = 0.0; ++found; }
                    if (r1 >= 3e-9) { v[i + blockDim.x] = 0.0; ++found; }
                }
            }
            if (s < S) {
                double* v = a + (size_t)s * n;
                const u64 key = ((u64)(salt * 64) << 40) + (u64)s * E;
                for (u64 x = tid; x < E; x += gs) atomicAdd(&v[target(x + key, n)], 1e-9);
            }
            grid.sync();
        }
        if (found == 0x7fffffffu) *sink = 1.0;
        return;
    }
    for (int s0 = (mode ? -K : 0); s0 < S; s0 += K) {
        if (mode) { // phase A of the next group
            for (int s = s0 + K; s < min(S, s0 + 2 * K); ++s) {
                double* v = a + (size_t)s * n;
                const u64 key = ((u64)(salt * 64 + 33) << 40) + (u64)s * F;
                for (u64 x = tid; x < F; x += gs) acc += __longlong_as_double(atomicExch((u64*)&v[target(x + key, n)], 0ull));
            }
        }
        for (int s = max(s0, 0); s < min(S, s0 + K) && s0 >= 0; ++s) {
            double* v = a + (size_t)s * n;
            const u64 key = ((u64)(salt * 64) << 40) + (u64)s * E;
            for (u64 x = tid; x < E; x += 2 * gs) {
                const u32 j0 = target(x + key, n), j1 = target(x + gs + key, n);
                const bool two = x + gs < E;
                const double o0 = atomicAdd(&v[j0], 1e-9);
                double o1 = 0;
                if (two) o1 = atomicAdd(&v[j1], 1e-9);
                acc += o0 + o1;
            }
        }
        grid.sync();
    }
    if (acc == 123.456) *sink = acc;
}
int main(int argc, char** argv) {
    u32 n = 4847571u;
    int S = atoi(argv[1]);
    u64 E = strtoull(argv[2], 0, 10);
    int K = atoi(argv[3]), mode = atoi(argv[4]);
    const int bps = argc > 5 ? atoi(argv[5]) : 2, threads = argc > 6 ? atoi(argv[6]) : 512;
    u64 F = E / 6;
    double *a, *sink;
    cudaMalloc(&a, (size_t)S * n * 8); cudaMalloc(&sink, 8);
    cudaMemset(a, 0, (size_t)S * n * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int G = 148 * bps, reps = 3;
    u32 salt = 0;
    void* args[] = {&a, &n, &S, &E, &F, &K, &mode, &salt, &sink};
    cudaLaunchCooperativeKernel((void*)sweep, dim3(G), dim3(threads), args, 0, 0);
    cudaEventRecord(e0);
    for (int r = 1; r <= reps; ++r) { salt = r; cudaLaunchCooperativeKernel((void*)sweep, dim3(G), dim3(threads), args, 0, 0); }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-6s S=%2d E=%8llu K=%d grid %dx%d: %7.2f G adds/s, %7.1f us per vector-level [%s]\n", mode == 0 ? "B-only" : mode == 1 ? "A|B" : mode == 2 ? "RED|scan" : "RED", S, E, K, G, threads,
           (double)S * E * reps / ms / 1e6, ms * 1e3 / reps / S, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
