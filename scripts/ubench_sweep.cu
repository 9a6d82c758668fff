// Microbenchmark behind the round-2 "slot-sequenced" push question (development aid, not product code):
// S dense fp64 vectors of n entries (the per-slot residue vectors, S * n * 8 bytes >> L2). One "level" scatters E
// random fp64 atomics with return into every vector. How fast is that as a function of the ORDER in which the grid
// visits the vectors and of the window of vectors that are live at the same time?
//   line   the engine's order: one slot-major line of S*E operations cut into tiles of T operations, tile t -> CTA t mod G
//   exch   a phase-A-like pass first (F = E/6 random atomicExch per vector over ALL vectors), then the line
//   fused  per vector: its F exchanges, then its E adds (the vector stays in L2 between the two)
//   pf     like line, but every CTA first streams its share of the vector (sequential 16-byte loads) when it enters a new vector
// Targets are uniform (skew 1) or skewed towards the low ids (idx = n * u^skew), which is what the in-degree relabelling gives.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_sweep scripts/ubench_sweep.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
typedef unsigned long long u64;
typedef unsigned int u32;
__device__ __forceinline__ u32 hash32(u32 x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ u32 target(u64 op, u32 n, int skew) {
    const u32 h = hash32((u32)op * 2654435761u + (u32)(op >> 32) * 40503u + 12345u);
    float u = (h >> 8) * (1.0f / 16777216.0f);
    float v = u;
    for (int k = 1; k < skew; ++k) v *= u;
    u32 j = (u32)(v * (float)n);
    return j < n ? j : n - 1;
}

// MODE 0 line, 1 fused (exchange part of a vector's tile range first), 2 pf
template <int MODE>
__global__ void __launch_bounds__(512, 2) sweep_kernel(double* a, u32 n, int S, u64 E, u64 T, int skew, u32 salt, double* sink) {
    const u64 total = (u64)S * E;
    double acc = 0;
    int last_s = -1;
    for (u64 lo = (u64)blockIdx.x * T; lo < total; lo += (u64)gridDim.x * T) {
        const u64 hi = lo + T < total ? lo + T : total;
        if (MODE == 2) {
            const int s = (int)(lo / E);
            if (s != last_s) { // stream this CTA's share of vector s into L2
                last_s = s;
                const double2* v = reinterpret_cast<const double2*>(a + (size_t)s * n);
                const u32 n2 = n / 2, per = (n2 + gridDim.x - 1) / gridDim.x;
                const u32 b = blockIdx.x * per, e = min(n2, b + per);
                for (u32 i = b + threadIdx.x; i < e; i += blockDim.x) {
                    double2 x;
                    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(x.x), "=d"(x.y) : "l"(v + i));
                    acc += x.x;
                }
            }
        }
        for (u64 x = lo + threadIdx.x; x < hi; x += 2 * blockDim.x) {
            const u64 x1 = x + blockDim.x;
            const int s0 = (int)(x / E), s1 = (int)(x1 / E);
            const u32 j0 = target(x + ((u64)salt << 40), n, skew), j1 = target(x1 + ((u64)salt << 40), n, skew);
            const double o0 = atomicAdd(&a[(size_t)s0 * n + j0], 1e-9);
            double o1 = 0;
            if (x1 < hi) o1 = atomicAdd(&a[(size_t)s1 * n + j1], 1e-9);
            acc += o0 + o1;
        }
    }
    if (acc == 123.456) *sink = acc;
}
__global__ void __launch_bounds__(512, 2) exch_kernel(double* a, u32 n, int S, u64 F, u64 T, u32 salt, double* sink) {
    const u64 total = (u64)S * F;
    double acc = 0;
    for (u64 lo = (u64)blockIdx.x * T; lo < total; lo += (u64)gridDim.x * T) {
        const u64 hi = lo + T < total ? lo + T : total;
        for (u64 x = lo + threadIdx.x; x < hi; x += blockDim.x) {
            const int s = (int)(x / F);
            const u32 j = target(x + ((u64)(salt + 77) << 40), n, 1);
            acc += __longlong_as_double(atomicExch((u64*)&a[(size_t)s * n + j], 0ull));
        }
    }
    if (acc == 123.456) *sink = acc;
}
// fused: vector by vector, every CTA does its share of the F exchanges and then its share of the E adds (no barrier in
// between: this measures the memory system, not the dependency)
__global__ void __launch_bounds__(512, 2) fused_kernel(double* a, u32 n, int S, u64 E, u64 F, int skew, u32 salt, double* sink) {
    double acc = 0;
    const u64 gs = (u64)gridDim.x * blockDim.x, tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    for (int s = 0; s < S; ++s) {
        double* v = a + (size_t)s * n;
        for (u64 x = tid; x < F; x += gs) acc += __longlong_as_double(atomicExch((u64*)&v[target(x + (u64)s * F + ((u64)(salt + 77) << 40), n, 1)], 0ull));
        for (u64 x = tid; x < E; x += 2 * gs) {
            const double o0 = atomicAdd(&v[target(x + (u64)s * E + ((u64)salt << 40), n, skew)], 1e-9);
            double o1 = 0;
            if (x + gs < E) o1 = atomicAdd(&v[target(x + gs + (u64)s * E + ((u64)salt << 40), n, skew)], 1e-9);
            acc += o0 + o1;
        }
    }
    if (acc == 123.456) *sink = acc;
}

int main(int argc, char** argv) {
    const u32 n = 4847571u;
    const int S = argc > 1 ? atoi(argv[1]) : 48;
    const int G = 296;
    double *a, *sink;
    cudaMalloc(&a, (size_t)S * n * 8); cudaMalloc(&sink, 8);
    cudaMemset(a, 0, (size_t)S * n * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int reps = 3;
    printf("# S=%d vectors of %u doubles (%.1f MB each); ops/s counts the E adds only\n", S, n, n * 8.0 / 1e6);
    for (int skew : {1, 3}) {
        for (u64 E : {500000ull, 1000000ull, 2000000ull, 4000000ull}) {
            const u64 F = E / 6;
            for (u64 T : {2048ull, 8192ull, 32768ull}) {
                for (int mode : {0, 2}) {
                    if (mode == 2 && T != 8192) continue;
                    if (mode == 0) sweep_kernel<0><<<G, 512>>>(a, n, S, E, T, skew, 0, sink);
                    else sweep_kernel<2><<<G, 512>>>(a, n, S, E, T, skew, 0, sink);
                    cudaEventRecord(e0);
                    for (int r = 1; r <= reps; ++r) {
                        if (mode == 0) sweep_kernel<0><<<G, 512>>>(a, n, S, E, T, skew, r, sink);
                        else sweep_kernel<2><<<G, 512>>>(a, n, S, E, T, skew, r, sink);
                    }
                    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
                    printf("%-5s skew %d E=%8llu T=%6llu: %7.2f G adds/s  (%.1f us per vector-level)\n", mode == 0 ? "line" : "pf", skew, E, T,
                           S * (double)E * reps / ms / 1e6, ms * 1e3 / reps / S);
                }
            }
            // phase-A-like pass over all vectors, then the line (two kernels = the barrier)
            cudaEventRecord(e0);
            for (int r = 1; r <= reps; ++r) {
                exch_kernel<<<G, 512>>>(a, n, S, F, 8192, r, sink);
                sweep_kernel<0><<<G, 512>>>(a, n, S, E, 8192, skew, r, sink);
            }
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("exch+line skew %d E=%8llu T=  8192: %7.2f G adds/s  (%.1f us per vector-level)\n", skew, E, S * (double)E * reps / ms / 1e6, ms * 1e3 / reps / S);
            cudaEventRecord(e0);
            for (int r = 1; r <= reps; ++r) exch_kernel<<<G, 512>>>(a, n, S, F, 8192, r, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("exch only        E=%8llu (F=%llu): %7.2f G exch/s  (%.1f us per vector-level)\n", E, F, S * (double)F * reps / ms / 1e6, ms * 1e3 / reps / S);
            cudaEventRecord(e0);
            for (int r = 1; r <= reps; ++r) fused_kernel<<<G, 512>>>(a, n, S, E, F, skew, r, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("fused skew %d E=%8llu          : %7.2f G adds/s  (%.1f us per vector-level)\n", skew, E, S * (double)E * reps / ms / 1e6, ms * 1e3 / reps / S);
        }
    }
    cudaError_t err = cudaDeviceSynchronize();
    printf("# %s\n", cudaGetErrorString(err));
    return 0;
}
