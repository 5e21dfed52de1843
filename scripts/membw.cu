// HBM microbenchmark for the K2 roofline: read-only, write-only, copy and K2-like 2-reads-1-write streams (256-bit accesses).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/bin/membw scripts/membw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
struct D4 { double x, y, z, w; };
__device__ __forceinline__ D4 ld(const double *p) { D4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p)); return v; }
__device__ __forceinline__ void st(double *p, D4 v) { asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory"); }
__global__ void k_read(const double *a, size_t n4, double *out) {
  double s = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) { D4 v = ld(a + i * 4); s += v.x + v.y + v.z + v.w; }
  if (s == 123.456) out[0] = s;
}
__global__ void k_write(double *a, size_t n4) {
  D4 v{1, 2, 3, 4};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) st(a + i * 4, v);
}
__global__ void k_copy(const double *a, double *b, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) st(b + i * 4, ld(a + i * 4));
}
__global__ void k_2r1w(const double *a, const double *b, double *c, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    D4 x = ld(a + i * 4), y = ld(b + i * 4);
    st(c + i * 4, D4{x.x * y.x, x.y * y.y, x.z * y.z, x.w * y.w});
  }
}
int main() {
  const size_t bytes = 4ull << 30, n4 = bytes / 32;
  double *a, *b, *c, *o;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&c, bytes); cudaMalloc(&o, 8);
  cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes); cudaMemset(c, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int blocks : {148 * 8, 148 * 16, 148 * 32}) for (int thr : {256, 512}) {
    float best[5] = {1e9, 1e9, 1e9, 1e9, 1e9};
    for (int rep = 0; rep < 6; ++rep) for (int k = 0; k < 5; ++k) {
      cudaEventRecord(e0);
      if (k == 0) k_read<<<blocks, thr>>>(a, n4, o);
      if (k == 1) k_write<<<blocks, thr>>>(a, n4);
      if (k == 2) k_copy<<<blocks, thr>>>(a, b, n4);
      if (k == 3) k_2r1w<<<blocks, thr>>>(a, b, c, n4);
      if (k == 4) cudaMemsetAsync(c, 1, bytes);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best[k]) best[k] = ms;
    }
    printf("blocks %5d x %3d: read %.0f GB/s  write %.0f GB/s  copy %.0f GB/s (r+w)  2r1w %.0f GB/s (r+r+w)  memset %.0f GB/s\n", blocks, thr,
           bytes / best[0] / 1e6, bytes / best[1] / 1e6, 2.0 * bytes / best[2] / 1e6, 3.0 * bytes / best[3] / 1e6, bytes / best[4] / 1e6);
  }
  return 0;
}
