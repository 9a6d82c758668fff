// Follow-up to ubench_sweep.cu (development aid): what does a FIRST touch of a cold sector cost, and are the later touches L2 hits?
// S vectors of n doubles; the grid visits the vectors in order, R consecutive visits of E random operations per vector
// (fresh random targets per visit). OP: 0 ATOM.f64 with return, 1 RED.f64, 2 ld.cg (read only), 3 ld + plain st (no atomic).
// usage: ubench_sweep2 S E R OP [blocks_per_sm] [threads]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
typedef unsigned long long u64;
typedef unsigned int u32;
__device__ __forceinline__ u32 hash32(u32 x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ u32 target(u64 op, u32 n) {
    const u32 h = hash32((u32)op * 2654435761u + (u32)(op >> 32) * 40503u + 12345u);
    return (u32)(((u64)h * n) >> 32);
}
template <int OP>
__global__ void sweep(double* a, u32 n, int S, u64 E, int R, u32 salt, double* sink) {
    double acc = 0;
    const u64 gs = (u64)gridDim.x * blockDim.x, tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    for (int s = 0; s < S; ++s) {
        double* v = a + (size_t)s * n;
        for (int r = 0; r < R; ++r) {
            const u64 key = ((u64)(salt * 64 + r) << 40) + (u64)s * E;
            for (u64 x = tid; x < E; x += 2 * gs) {
                const u32 j0 = target(x + key, n), j1 = target(x + gs + key, n);
                const bool two = x + gs < E;
                if (OP == 0) { const double o0 = atomicAdd(&v[j0], 1e-9); double o1 = 0; if (two) o1 = atomicAdd(&v[j1], 1e-9); acc += o0 + o1; }
                if (OP == 1) { atomicAdd(&v[j0], 1e-9); if (two) atomicAdd(&v[j1], 1e-9); }
                if (OP == 2) { const double o0 = __ldcg(&v[j0]); double o1 = 0; if (two) o1 = __ldcg(&v[j1]); acc += o0 + o1; }
                if (OP == 3) { const double o0 = __ldcg(&v[j0]); double o1 = 0; if (two) o1 = __ldcg(&v[j1]); __stcg(&v[j0], o0 + 1e-9); if (two) __stcg(&v[j1], o1 + 1e-9); }
            }
        }
    }
    if (acc == 123.456) *sink = acc;
}
int main(int argc, char** argv) {
    const u32 n = 4847571u;
    const int S = atoi(argv[1]);
    const u64 E = strtoull(argv[2], 0, 10);
    const int R = atoi(argv[3]), OP = atoi(argv[4]);
    const int bps = argc > 5 ? atoi(argv[5]) : 2, threads = argc > 6 ? atoi(argv[6]) : 512;
    double *a, *sink;
    cudaMalloc(&a, (size_t)S * n * 8); cudaMalloc(&sink, 8);
    cudaMemset(a, 0, (size_t)S * n * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int G = 148 * bps, reps = 3;
    auto launch = [&](u32 salt) {
        if (OP == 0) sweep<0><<<G, threads>>>(a, n, S, E, R, salt, sink);
        if (OP == 1) sweep<1><<<G, threads>>>(a, n, S, E, R, salt, sink);
        if (OP == 2) sweep<2><<<G, threads>>>(a, n, S, E, R, salt, sink);
        if (OP == 3) sweep<3><<<G, threads>>>(a, n, S, E, R, salt, sink);
    };
    launch(0);
    cudaEventRecord(e0);
    for (int r = 1; r <= reps; ++r) launch(r);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    static const char* names[] = {"ATOM", "RED", "LD", "LD+ST"};
    printf("%-5s S=%2d E=%8llu R=%d grid %dx%d: %7.2f G ops/s, %7.1f us per vector (all %d visits) [%s]\n", names[OP], S, E, R, G, threads,
           (double)S * E * R * reps / ms / 1e6, ms * 1e3 / reps / S, R, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
