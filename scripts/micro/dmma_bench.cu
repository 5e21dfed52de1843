// FP64 tensor-core (DMMA) throughput per mma.sync shape on sm_100a: which shape should the 20-state kernels issue?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/dmma_bench scripts/micro/dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE, int CHAINS>
__global__ void __launch_bounds__(128) k(double *out, int iters, double a0, double b0) {
  double acc[CHAINS][4];
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) acc[c][i] = 0.0;
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = a0 + threadIdx.x * 1e-9 + i;
  for (int i = 0; i < 4; ++i) b[i] = b0 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (SHAPE == 0) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[c][0]), "+d"(acc[c][1]) : "d"(a[0]), "d"(b[0]));
      if (SHAPE == 1) asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};" : "+d"(acc[c][0]), "+d"(acc[c][1]), "+d"(acc[c][2]), "+d"(acc[c][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
      if (SHAPE == 2) asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+d"(acc[c][0]), "+d"(acc[c][1]), "+d"(acc[c][2]), "+d"(acc[c][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
      if (SHAPE == 3) asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};" : "+d"(acc[c][0]), "+d"(acc[c][1]), "+d"(acc[c][2]), "+d"(acc[c][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
  }
  double s = 0;
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) s += acc[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// plain FP64 FMA pipe for comparison
template <int CHAINS>
__global__ void __launch_bounds__(128) kf(double *out, int iters, double a0, double b0) {
  double acc[CHAINS];
  for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x + c;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a0, b0);
  }
  double s = 0;
  for (int c = 0; c < CHAINS; ++c) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  double *out; cudaMalloc(&out, 148 * 16 * 128 * sizeof(double));
  const int iters = 20000, blocks = 148 * 8;
  const double flops[4] = {8. * 8 * 4 * 2, 16. * 8 * 4 * 2, 16. * 8 * 8 * 2, 16. * 8 * 16 * 2};
  const char *names[4] = {"m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16"};
  for (int warps = 1; warps <= 4; warps *= 2) {
    float ms;
    ms = timeit([&] { k<0, 6><<<blocks, 32 * warps>>>(out, iters, 1.0, 2.0); });
    printf("%-9s warps/block %d chains 6: %.2f TFLOP/s\n", names[0], warps, flops[0] * 6 * iters * blocks * warps / ms / 1e9);
    ms = timeit([&] { k<1, 6><<<blocks, 32 * warps>>>(out, iters, 1.0, 2.0); });
    printf("%-9s warps/block %d chains 6: %.2f TFLOP/s\n", names[1], warps, flops[1] * 6 * iters * blocks * warps / ms / 1e9);
    ms = timeit([&] { k<2, 6><<<blocks, 32 * warps>>>(out, iters, 1.0, 2.0); });
    printf("%-9s warps/block %d chains 6: %.2f TFLOP/s\n", names[2], warps, flops[2] * 6 * iters * blocks * warps / ms / 1e9);
    ms = timeit([&] { k<3, 6><<<blocks, 32 * warps>>>(out, iters, 1.0, 2.0); });
    printf("%-9s warps/block %d chains 6: %.2f TFLOP/s\n", names[3], warps, flops[3] * 6 * iters * blocks * warps / ms / 1e9);
    ms = timeit([&] { kf<8><<<blocks, 32 * warps>>>(out, iters * 8, 1.0000001, 1e-9); });
    printf("%-9s warps/block %d chains 8: %.2f TFLOP/s\n", "DFMA", warps, 2.0 * 8 * iters * 8 * blocks * warps * 32 / ms / 1e9);
  }
  // how many resident warps per SM does the FP64 tensor path need?  148 x W one-warp blocks (W warps per SM)
  for (int W : {1, 2, 3, 4, 6, 8, 12, 16}) {
    float ms = timeit([&] { k<0, 6><<<148 * W, 32>>>(out, iters, 1.0, 2.0); });
    float ms12 = timeit([&] { k<0, 12><<<148 * W, 32>>>(out, iters, 1.0, 2.0); });
    float ms3 = timeit([&] { k<0, 3><<<148 * W, 32>>>(out, iters, 1.0, 2.0); });
    printf("m8n8k4 %2d warps/SM: 3 chains %.2f, 6 chains %.2f, 12 chains %.2f TFLOP/s\n", W, flops[0] * 3 * iters * 148 * W / ms3 / 1e9,
           flops[0] * 6 * iters * 148 * W / ms / 1e9, flops[0] * 12 * iters * 148 * W / ms12 / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
