#include <string.h>
// Microbenchmark (development aid): throughput of random fp64 atomics on B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ uint32_t hash32(uint32_t x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }
template<int MODE> __global__ void k(double* a, const int* deg, uint32_t n, u64 total, double* sink, unsigned* ctr){
  u64 tid = blockIdx.x*(u64)blockDim.x+threadIdx.x, gs = gridDim.x*(u64)blockDim.x; double acc=0;
  for(u64 i=tid;i<total;i+=gs){ uint32_t j = (uint32_t)(((u64)hash32((uint32_t)i*2654435761u+12345u)*n)>>32);
    if(MODE==0) acc += atomicAdd(&a[j], 1e-9);            // ATOM f64 with return
    else if(MODE==1) atomicAdd(&a[j], 1e-9);               // RED f64
    else if(MODE==2) acc += __ldg(&a[j]);                  // random 8B load
    else if(MODE==3) { double o=atomicAdd(&a[j],1e-9); int d=__ldg(&deg[j]); acc += (o < 1e-3*d); } // ATOM + deg load
    else if(MODE==4) { acc += __longlong_as_double(atomicAdd((u64*)&a[j], 1ull)); } // u64 ATOM
    else if(MODE==5) { atomicAdd((float*)&a[j], 1e-9f); } // f32 RED
    else if(MODE==6) { acc += atomicAdd((float*)&a[j], 1e-9f); } // f32 ATOM
    else if(MODE==7) { atomicAdd(ctr, 1u); } // same-address RED u32
    else if(MODE==8) { acc += atomicAdd(ctr, 1u); } // same-address ATOM u32 (per thread)
    else if(MODE==9) { acc += __ldg(&((const int*)a)[j]); } // random 4B load (n counts ints)
  }
  if(acc==123.456) *sink=acc;
}
template<int MODE> void run(const char* name, double* a, int* deg, uint32_t n, u64 total, double* sink, unsigned* ctr, int grid, int block){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid,block>>>(a,deg,n,total/8,sink,ctr); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<grid,block>>>(a,deg,n,total,sink,ctr); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1); printf("%-28s n=%9u (%6.1f MB) grid=%d x %d : %8.2f G ops/s\n", name, n, n*8.0/1e6, grid, block, total/ms/1e6);
}
int main(int argc, char** argv){ double* a; int* deg; double* sink; unsigned* ctr; uint32_t nmax=80000000; cudaMalloc(&a, nmax*8ull); cudaMalloc(&deg,nmax*4ull); cudaMalloc(&sink,8); cudaMalloc(&ctr,4);
  cudaMemset(a,0,nmax*8ull); cudaMemset(deg,1,nmax*4ull); cudaMemset(ctr,0,4);
  u64 total = 1ull<<30;
  // argv[1] = "gran": only the L2 fill-granularity comparison (cudaLimitMaxL2FetchGranularity 32 / 64 / 128 bytes)
  if (argc > 1 && !strcmp(argv[1], "rand4")) { // random 4-byte reads: the neighbour fetch of a walk without the dependency chain
    for (uint32_t n : {16u<<20, 32u<<20, 69000000u, 138000000u})
      for (int grid : {148*8, 148*16}) run<9>("LDG.32 random (n ints)",a,deg,n,total,sink,ctr,grid,256);
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "gran")) {
    for (size_t gran : {(size_t)64, (size_t)32, (size_t)128, (size_t)64}) {
      cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
      printf("== cudaLimitMaxL2FetchGranularity set %zu -> %s, now %zu\n", gran, cudaGetErrorString(e), got);
      for (uint32_t n : {4847571u*4, 4847571u*16}) {
        run<0>("ATOM.f64 (return)",a,deg,n,total,sink,ctr,296,512);
        run<2>("LDG.64 random",a,deg,n,total,sink,ctr,296,512);
        run<3>("ATOM.f64 + LDG deg",a,deg,n,total,sink,ctr,296,512);
      }
    }
    return 0;
  }
  for (uint32_t n : {4847571u, 4847571u*4, 4847571u*16}) {
    for (int grid : {148*2, 148*8}) { int block = grid==296?512:256;
      run<0>("ATOM.f64 (return)",a,deg,n,total,sink,ctr,grid,block);
      run<1>("RED.f64",a,deg,n,total,sink,ctr,grid,block);
      run<2>("LDG.64 random",a,deg,n,total,sink,ctr,grid,block);
      run<3>("ATOM.f64 + LDG deg",a,deg,n,total,sink,ctr,grid,block);
      run<4>("ATOM.u64",a,deg,n,total,sink,ctr,grid,block);
      run<5>("RED.f32",a,deg,n,total,sink,ctr,grid,block);
      run<6>("ATOM.f32",a,deg,n,total,sink,ctr,grid,block);
    }
  }
  run<7>("RED.u32 same address",a,deg,4847571u,1ull<<26,sink,ctr,296,512);
  run<8>("ATOM.u32 same address",a,deg,4847571u,1ull<<24,sink,ctr,296,512);
  return 0; }
