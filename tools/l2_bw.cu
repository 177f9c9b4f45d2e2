// L2-resident read bandwidth of the GPU: every CTA re-reads a buffer that fits the L2 `reps` times
// with 256-bit loads.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/l2_bw.cu -o /tmp/l2_bw
#include <cstdio>
#include <cuda_runtime.h>
struct __align__(32) d4 { double x, y, z, w; };
__global__ void k_read(const d4* __restrict__ p, size_t n, int reps, double* out) {
  double acc = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      d4 v;
      asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p + i));
      acc += v.x + v.y + v.z + v.w;
    }
  if (acc == 123.456) *out = acc;
}
int main() {
  for (size_t mb : {16, 32, 64, 96, 2048}) {
    const size_t n = mb * 1024 * 1024 / sizeof(d4);
    d4* p; double* o;
    cudaMalloc(&p, n * sizeof(d4)); cudaMalloc(&o, 8); cudaMemset(p, 0, n * sizeof(d4));
    const int reps = mb > 1000 ? 4 : 200;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads : {256, 512, 1024}) {
      k_read<<<148 * (2048 / threads), threads>>>(p, n, 2, o);
      cudaEventRecord(e0);
      k_read<<<148 * (2048 / threads), threads>>>(p, n, reps, o);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("%zu MB, %d threads/CTA: %.0f GB/s\n", mb, threads, (double)mb / 1024 * reps / (ms * 1e-3));
    }
    cudaFree(p); cudaFree(o);
  }
  return 0;
}
