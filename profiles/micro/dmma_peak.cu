// Microbenchmark (profiles only, not part of the library): fp64 mma.sync.m8n8k4 throughput and dependent-issue
// latency on this GPU next to plain DFMA.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a dmma_peak.cu -o dmma_peak
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CHAINS>
__global__ void dmma_kernel(double *sink, int iters, double a, double b) {
  double c[CHAINS][2];
  for (int k = 0; k < CHAINS; ++k) { c[k][0] = threadIdx.x; c[k][1] = k; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) dmma(c[k][0], c[k][1], a, b);
  }
  double s = 0;
  for (int k = 0; k < CHAINS; ++k) s += c[k][0] + c[k][1];
  if (s == 12345.678) *sink = s;
}
template <int CHAINS>
__global__ void dfma_kernel(double *sink, int iters, double a, double b) {
  double c[CHAINS];
  for (int k = 0; k < CHAINS; ++k) c[k] = threadIdx.x + k;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) c[k] = fma(c[k], a, b);
  }
  double s = 0;
  for (int k = 0; k < CHAINS; ++k) s += c[k];
  if (s == 12345.678) *sink = s;
}
template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double *sink; cudaMalloc(&sink, 8);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 1 << 14;
  for (int warps = 1; warps <= 16; warps *= 2) {
    const int grid = sms, threads = 32 * warps * 4;   // `warps` warps per scheduler
    float m1 = timeit([&] { dmma_kernel<1><<<grid, threads>>>(sink, iters, 0.999, 1e-9); });
    float m4 = timeit([&] { dmma_kernel<4><<<grid, threads>>>(sink, iters, 0.999, 1e-9); });
    float m8 = timeit([&] { dmma_kernel<8><<<grid, threads>>>(sink, iters, 0.999, 1e-9); });
    float f8 = timeit([&] { dfma_kernel<8><<<grid, threads>>>(sink, iters, 0.999, 1e-9); });
    const double wi = (double)grid * threads / 32 * iters;
    printf("warps/sched %2d: DMMA 1 chain %.2f TFLOP/s (%.1f clk/inst/warp), 4 chains %.2f, 8 chains %.2f | DFMA 8 chains %.2f TFLOP/s\n", warps,
           wi * 512 / (m1 * 1e-3) / 1e12, m1 * 1e-3 * 1.965e9 / iters, wi * 4 * 512 / (m4 * 1e-3) / 1e12, wi * 8 * 512 / (m8 * 1e-3) / 1e12,
           wi * 8 * 64 / (f8 * 1e-3) / 1e12);
  }
  return 0;
}
