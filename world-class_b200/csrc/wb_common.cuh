// Common host/device helpers for the worldb200 CUDA path (sm_100a).
//
// All arithmetic that feeds an integer decision (rounding to a sample index, window
// half-lengths, voiced/unvoiced thresholds) is written so that nvcc cannot contract it
// into FMAs: the whole library is compiled with --fmad=false and FMAs are only used
// where they are spelled out with fma() (FFT twiddle products, dot products).  The
// reference is built with g++ -O3 -mavx, i.e. without FMA contraction
// (/root/reference/Makefile:13).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define WB_PI 3.1415926535897932384          // world::kPi, world_constantnumbers.hpp:12
#define WB_SAFEGUARD 0.000000000001          // world::kMySafeGuardMinimum
#define WB_EPS 0.00000000000000022204460492503131  // world::kEps
#define WB_DEFAULT_F0 500.0                  // world::kDefaultF0
#define WB_LOG2 0.69314718055994529          // world::kLog2
#define WB_FLOOR_F0_D4C 47.0                 // world::kFloorF0D4C
#define WB_FREQ_INTERVAL 3000.0              // world::kFrequencyInterval
#define WB_UPPER_LIMIT 15000.0               // world::kUpperLimit

#define WB_OK 0
#define WB_ERR_CUDA 1
#define WB_ERR_ARG 2
#define WB_ERR_UNSUPPORTED 3

typedef double2 cplx;

#define WB_CUDA_CHECK(expr)                                                         \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      fprintf(stderr, "worldb200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e), \
              __FILE__, __LINE__, cudaGetErrorString(_e));                          \
      return WB_ERR_CUDA;                                                           \
    }                                                                               \
  } while (0)

// matlab_round(): half away from zero via truncation (world_matlabfunctions.cpp:212-214)
__host__ __device__ __forceinline__ int wb_round(double x) {
  return x > 0 ? static_cast<int>(x + 0.5) : static_cast<int>(x - 0.5);
}
__host__ __device__ __forceinline__ int wb_min_i(int a, int b) { return a < b ? a : b; }
__host__ __device__ __forceinline__ int wb_max_i(int a, int b) { return a > b ? a : b; }

// ---- block-wide reductions / scans (blockDim.x multiple of 32, <= 1024) -----------
__device__ __forceinline__ double wb_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the block; `red` is a shared array of >= 32 doubles.  All threads get the
// result.  Contains two __syncthreads().
__device__ __forceinline__ double wb_block_sum(double v, double *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = wb_warp_sum(v);
  __syncthreads();  // protect `red` from a previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  double t = (lane < nw) ? red[lane] : 0.0;
  t = wb_warp_sum(t);
  return t;
}

// Two sums at once.
__device__ __forceinline__ void wb_block_sum2(double &a, double &b, double *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a = wb_warp_sum(a);
  b = wb_warp_sum(b);
  __syncthreads();
  if (lane == 0) { red[warp] = a; red[32 + warp] = b; }
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  double ta = (lane < nw) ? red[lane] : 0.0;
  double tb = (lane < nw) ? red[32 + lane] : 0.0;
  a = wb_warp_sum(ta);
  b = wb_warp_sum(tb);
}

// In-place inclusive prefix sum of a shared array a[0..n).  Each thread owns a
// contiguous chunk (sequential inside the chunk, like the reference's running sum),
// chunk totals are scanned through `red` (>= blockDim.x doubles... we use 1024).
__device__ inline void wb_block_inclusive_scan(double *a, int n, double *red) {
  const int nt = blockDim.x, tid = threadIdx.x;
  const int chunk = (n + nt - 1) / nt;
  const int b = tid * chunk;
  const int e = min(n, b + chunk);
  double s = 0.0;
  for (int i = b; i < e; ++i) { s += a[i]; a[i] = s; }
  __syncthreads();
  red[tid] = s;
  __syncthreads();
  if (tid < 32) {
    // serial-by-lane scan over nt partials (nt <= 1024 -> <= 32 per lane)
    const int per = (nt + 31) / 32;
    const int pb = tid * per, pe = min(nt, pb + per);
    double acc = 0.0;
    for (int i = pb; i < pe; ++i) { acc += red[i]; red[i] = acc; }
    double incl = acc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double t = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += t;
    }
    const double excl = incl - acc;
    for (int i = pb; i < pe; ++i) red[i] += excl;
  }
  __syncthreads();
  if (tid > 0) {
    const double off = red[tid - 1];
    for (int i = b; i < e; ++i) a[i] += off;
  }
  __syncthreads();
}

// Single-CTA exclusive prefix sum over the per-item counts count(i), i in [0, n): offsets[i], offsets[n] = total;
// optionally publishes *skip_out = *skip_in + total (randn stream bookkeeping, see WbRngCursor).
// count(i) is evaluated twice (it is a cheap closed form everywhere).  blockDim.x multiple of 32, <= 1024.
template <typename F>
__device__ __forceinline__ void wb_block_count_scan(F count, int n, unsigned long long *__restrict__ offsets,
                                                    const unsigned long long *__restrict__ skip_in,
                                                    unsigned long long *__restrict__ skip_out) {
  __shared__ unsigned long long s_warp_total[32];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const int chunk = (n + nt - 1) / nt;
  const int b = min(n, tid * chunk), e = min(n, b + chunk);
  unsigned long long s = 0;
  for (int i = b; i < e; ++i) s += count(i);
  unsigned long long incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp_total[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = (lane < nw) ? s_warp_total[lane] : 0ull;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    s_warp_total[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  unsigned long long run = incl - s + (warp > 0 ? s_warp_total[warp - 1] : 0ull);
  for (int i = b; i < e; ++i) { offsets[i] = run; run += count(i); }
  if (tid == 0) {
    const unsigned long long total = s_warp_total[nw - 1];
    offsets[n] = total;
    if (skip_out) *skip_out = (skip_in ? *skip_in : 0ull) + total;
  }
}

// ---- launch accounting / per-kernel timing (wb_runtime.cu) ---------------------------------
// Every kernel launch of the library goes through WB_LAUNCH: it counts launches (bench.py's
// gpu_launches) and, when profiling is enabled, brackets the launch with CUDA events on the
// launching stream so that bench.py can report the dominant kernel's live duration.
struct WbLaunchScope {
  WbLaunchScope(const char *name, cudaStream_t stream);
  ~WbLaunchScope();
  const char *name_;
  cudaStream_t stream_;
  void *ev0_;
};
#define WB_LAUNCH(NAME, ...)               \
  do {                                     \
    WbLaunchScope _wb_scope(NAME, stream); \
    __VA_ARGS__;                           \
  } while (0)
unsigned long long wb_launch_counter();
void wb_launch_counter_add(unsigned long long n);  // kernels replayed through a CUDA graph
int wb_prof_is_enabled();
void wb_prof_set_enabled(int on);
// synchronises the device, folds all pending event pairs into per-name totals
int wb_prof_collect();
int wb_prof_query(const char *name, double *total_ms, int *count);
int wb_prof_names(char *buf, int buf_len);  // ';'-separated
void wb_prof_reset();
