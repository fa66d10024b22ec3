// Host runtime: workspaces, twiddle tables, small utility kernels.
#include "wb_internal.h"
#include "wb_scan.cuh"

#include <math.h>
#include <string.h>
#include <mutex>
#include <vector>

WbWorkspace::WbWorkspace() : generation_(0), d_err_(nullptr) {}

WbWorkspace::~WbWorkspace() {
  for (auto &kv : dev_) cudaFree(kv.second.p);
  for (auto &kv : pinned_) cudaFreeHost(kv.second.p);
  if (d_err_) cudaFreeHost(d_err_);
}

void *WbWorkspace::get(const std::string &name, size_t bytes) {
  if (bytes == 0) bytes = 8;
  auto it = dev_.find(name);
  if (it != dev_.end() && it->second.bytes >= bytes) return it->second.p;
  if (it != dev_.end()) { cudaFree(it->second.p); dev_.erase(it); ++generation_; }
  void *p = nullptr;
  const size_t grow = bytes + bytes / 8;  // a little slack so slowly growing inputs do not realloc every call
  if (cudaMalloc(&p, grow) != cudaSuccess) {
    fprintf(stderr, "worldb200: cudaMalloc(%zu) failed for %s\n", grow, name.c_str());
    return nullptr;
  }
  dev_[name] = Buf{p, grow};
  return p;
}

void *WbWorkspace::get_keep(const std::string &name, size_t bytes, size_t keep_bytes, cudaStream_t stream) {
  auto it = dev_.find(name);
  if (it == dev_.end() || keep_bytes == 0 || it->second.bytes >= bytes) return get(name, bytes);
  void *p = nullptr;
  const size_t grow = 2 * bytes;   // geometric: an ever-growing history reallocates O(log n) times
  if (cudaMalloc(&p, grow) != cudaSuccess) {
    fprintf(stderr, "worldb200: cudaMalloc(%zu) failed for %s\n", grow, name.c_str());
    return nullptr;
  }
  const size_t keep = keep_bytes < it->second.bytes ? keep_bytes : it->second.bytes;
  if (cudaMemcpyAsync(p, it->second.p, keep, cudaMemcpyDeviceToDevice, stream) != cudaSuccess ||
      cudaStreamSynchronize(stream) != cudaSuccess) {
    cudaFree(p);
    return nullptr;
  }
  cudaFree(it->second.p);
  ++generation_;
  it->second = Buf{p, grow};
  return p;
}

void *WbWorkspace::find(const std::string &name, size_t *bytes_out) const {
  auto it = dev_.find(name);
  if (it == dev_.end()) { if (bytes_out) *bytes_out = 0; return nullptr; }
  if (bytes_out) *bytes_out = it->second.bytes;
  return it->second.p;
}

void *WbWorkspace::get_pinned(const std::string &name, size_t bytes) {
  if (bytes == 0) bytes = 8;
  auto it = pinned_.find(name);
  if (it != pinned_.end() && it->second.bytes >= bytes) return it->second.p;
  if (it != pinned_.end()) { cudaFreeHost(it->second.p); pinned_.erase(it); }
  void *p = nullptr;
  const size_t grow = bytes + bytes / 8;
  if (cudaMallocHost(&p, grow) != cudaSuccess) {
    fprintf(stderr, "worldb200: cudaMallocHost(%zu) failed for %s\n", grow, name.c_str());
    return nullptr;
  }
  memset(p, 0, grow);  // (cache tags kept in pinned buffers rely on a zero start)
  pinned_[name] = Buf{p, grow};
  return p;
}

// The flag lives in page-locked host memory mapped into the device's address space (the same pointer on both sides,
// UVA): kernels set it on their -- rare -- error paths, and the host reads it after a stream synchronisation
// without another device-to-host copy (the class API checks it at the end of every compute() call).
int *WbWorkspace::error_flag() {
  if (!d_err_) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return nullptr;
    d_err_ = static_cast<int *>(p);
    *d_err_ = 0;
  }
  return d_err_;
}

int WbWorkspace::read_error_flag(cudaStream_t stream) {
  if (!d_err_) return WB_OK;
  if (cudaStreamSynchronize(stream) != cudaSuccess) return WB_ERR_CUDA;
  volatile int *flag = d_err_;
  const int h = *flag;
  if (h != 0) *flag = 0;
  return h;
}

namespace {
std::mutex g_tw_mutex;
std::map<int, cplx *> g_tw;

#define SCAN_THREADS 1024
__global__ void __launch_bounds__(SCAN_THREADS) scan_u64_kernel(const unsigned long long *__restrict__ counts,
                                                                unsigned long long *__restrict__ offsets, int n,
                                                                const unsigned long long *__restrict__ skip_in,
                                                                unsigned long long *__restrict__ skip_out) {
  wb_block_count_scan([&](int i) { return counts[i]; }, n, offsets, skip_in, skip_out);
}
}  // namespace

int wb_sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
    else sms = 148;   // B200
  }
  return sms;
}

const cplx *wb_twiddle_table(int n) {
  std::lock_guard<std::mutex> lock(g_tw_mutex);
  auto it = g_tw.find(n);
  if (it != g_tw.end()) return it->second;
  std::vector<cplx> h(n);
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int k = 0; k < n; ++k) {
    // exact symmetries first so that e.g. T[n/4] = (0, 1) exactly
    const int oct = (int)(((long long)k * 8) / n);
    (void)oct;
    const long double a = two_pi * (long double)k / (long double)n;
    h[k].x = (double)cosl(a);
    h[k].y = (double)sinl(a);
  }
  if (n >= 4) {
    h[0] = make_double2(1.0, 0.0);
    h[n / 4] = make_double2(0.0, 1.0);
    h[n / 2] = make_double2(-1.0, 0.0);
    h[3 * n / 4] = make_double2(0.0, -1.0);
  }
  cplx *d = nullptr;
  if (cudaMalloc(&d, sizeof(cplx) * n) != cudaSuccess) return nullptr;
  if (cudaMemcpy(d, h.data(), sizeof(cplx) * n, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  g_tw[n] = d;
  return d;
}

namespace {
struct CountsFn {
  const unsigned long long *counts;
  __device__ unsigned long long operator()(int i) const { return counts[i]; }
};

// one CTA: tile_off <- exclusive scan of tile_sum, stream bookkeeping (see wb_scan.cuh)
__global__ void __launch_bounds__(SCAN_THREADS) tile_scan_kernel(const unsigned long long *__restrict__ tile_sum, int n_tiles,
                                                                 unsigned long long *__restrict__ tile_off,
                                                                 unsigned long long *__restrict__ total_out,
                                                                 const unsigned long long *__restrict__ skip_in,
                                                                 const unsigned long long *__restrict__ skip_add,
                                                                 unsigned long long *__restrict__ skip_mid,
                                                                 unsigned long long *__restrict__ skip_out) {
  wb_block_count_scan([&](int i) { return tile_sum[i]; }, n_tiles, tile_off, nullptr, nullptr);
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long total = tile_off[n_tiles];
    const unsigned long long base = (skip_in ? *skip_in : 0ull) + (skip_add ? *skip_add : 0ull);
    *total_out = total;
    if (skip_mid) *skip_mid = base;
    if (skip_out) *skip_out = base + total;
  }
}

__global__ void tile_add_kernel(unsigned long long *__restrict__ offsets, int n, const unsigned long long *__restrict__ tile_off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) offsets[i] += tile_off[i / WB_SCAN_TILE];
}
}  // namespace

int wb_tile_scan_finish(unsigned long long *d_offsets, int n, unsigned long long *d_tile_sum, unsigned long long *d_tile_off,
                        int n_tiles, const unsigned long long *d_skip_in, const unsigned long long *d_skip_add,
                        unsigned long long *d_skip_mid, unsigned long long *d_skip_out, cudaStream_t stream) {
  WB_LAUNCH("tile_scan_kernel", tile_scan_kernel<<<1, SCAN_THREADS, 0, stream>>>(d_tile_sum, n_tiles, d_tile_off, d_offsets + n, d_skip_in,
                                                                              d_skip_add, d_skip_mid, d_skip_out));
  WB_LAUNCH("tile_add_kernel", tile_add_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_offsets, n, d_tile_off));
  WB_CUDA_CHECK(cudaGetLastError());
  return WB_OK;
}

int wb_exclusive_scan_u64(const unsigned long long *d_counts, unsigned long long *d_offsets, int n,
                          cudaStream_t stream, const unsigned long long *d_skip_in, unsigned long long *d_skip_out,
                          WbWorkspace *ws) {
  if (ws && n > WB_SCAN_SINGLE_CTA_MAX) {   // long streams: the whole GPU instead of one CTA
    CountsFn fn = {d_counts};
    return wb_count_scan_tiles(fn, n, d_offsets, d_skip_in, nullptr, nullptr, d_skip_out, ws, "scan_u64_tiles", stream);
  }
  WB_LAUNCH("scan_u64_kernel", scan_u64_kernel<<<1, SCAN_THREADS, 0, stream>>>(d_counts, d_offsets, n, d_skip_in, d_skip_out));
  WB_CUDA_CHECK(cudaGetLastError());
  return WB_OK;
}

// ---- launch accounting / per-kernel timing -------------------------------------------------
#include <atomic>
#include <string>
namespace {
std::atomic<unsigned long long> g_launches{0};
std::atomic<int> g_prof_on{0};
struct Pending { std::string name; cudaEvent_t e0, e1; };
std::mutex g_prof_mutex;
std::vector<Pending> g_pending;
struct Tot { double ms; int count; };
std::map<std::string, Tot> g_totals;
}  // namespace

WbLaunchScope::WbLaunchScope(const char *name, cudaStream_t stream) : name_(name), stream_(stream), ev0_(nullptr) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (g_prof_on.load(std::memory_order_relaxed)) {
    cudaEvent_t e0;
    if (cudaEventCreate(&e0) == cudaSuccess) {
      cudaEventRecord(e0, stream_);
      ev0_ = e0;
    }
  }
}

WbLaunchScope::~WbLaunchScope() {
  if (ev0_) {
    cudaEvent_t e1;
    if (cudaEventCreate(&e1) == cudaSuccess) {
      cudaEventRecord(e1, stream_);
      std::lock_guard<std::mutex> lock(g_prof_mutex);
      g_pending.push_back(Pending{name_, (cudaEvent_t)ev0_, e1});
    }
  }
}

namespace {
__global__ void range_offsets_kernel(const unsigned long long *__restrict__ offsets, int begin, int end,
                                     unsigned long long *__restrict__ rel, const unsigned long long *__restrict__ skip_in,
                                     unsigned long long *__restrict__ skip_range, unsigned long long *__restrict__ count_range) {
  const unsigned long long base = offsets[begin];
  const int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= end) rel[i] = offsets[i] - base;
  if (i == begin) {
    *skip_range = (skip_in ? *skip_in : 0ull) + base;
    *count_range = offsets[end] - base;
  }
}
}  // namespace

int wb_range_offsets(const unsigned long long *d_offsets, WbFrameRange range, unsigned long long *d_rel,
                     const unsigned long long *d_skip_in, unsigned long long *d_skip_range,
                     unsigned long long *d_count_range, cudaStream_t stream) {
  const int n = range.end - range.begin + 1;
  WB_LAUNCH("range_offsets_kernel", range_offsets_kernel<<<(n + 255) / 256, 256, 0, stream>>>(
      d_offsets, range.begin, range.end, d_rel, d_skip_in, d_skip_range, d_count_range));
  WB_CUDA_CHECK(cudaGetLastError());
  return WB_OK;
}

unsigned long long wb_launch_counter() { return g_launches.load(); }
void wb_launch_counter_add(unsigned long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int wb_prof_is_enabled() { return g_prof_on.load(std::memory_order_relaxed); }
void wb_prof_set_enabled(int on) { g_prof_on.store(on ? 1 : 0); }

int wb_prof_collect() {
  if (cudaDeviceSynchronize() != cudaSuccess) return WB_ERR_CUDA;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  for (auto &p : g_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
      Tot &t = g_totals[p.name];
      t.ms += ms;
      t.count += 1;
    }
    cudaEventDestroy(p.e0);
    cudaEventDestroy(p.e1);
  }
  g_pending.clear();
  return WB_OK;
}

int wb_prof_query(const char *name, double *total_ms, int *count) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  auto it = g_totals.find(name);
  if (it == g_totals.end()) return WB_ERR_ARG;
  *total_ms = it->second.ms;
  *count = it->second.count;
  return WB_OK;
}

int wb_prof_names(char *buf, int buf_len) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  std::string s;
  for (auto &kv : g_totals) { if (!s.empty()) s += ";"; s += kv.first; }
  if ((int)s.size() + 1 > buf_len) return WB_ERR_ARG;
  memcpy(buf, s.c_str(), s.size() + 1);
  return WB_OK;
}

void wb_prof_reset() {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  g_totals.clear();
}

// ---- fp64 multiply-add peak of this GPU (bench.py: the second roofline next to HBM bandwidth) ----
namespace {
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *__restrict__ sink, int iters, double a, double b) {
  // eight independent chains per thread: enough to cover the fp64 pipe latency at 8 warps per scheduler
  double x0 = threadIdx.x, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 0.123456789) sink[0] = s;   // keeps the chains alive; practically never taken
}
}  // namespace

int wb_measure_fp64_peak_tflops(double *tflops_out) {
  if (!tflops_out) return WB_ERR_ARG;
  cudaStream_t stream = nullptr;
  double *d_sink = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (cudaMalloc(&d_sink, sizeof(double)) != cudaSuccess) return WB_ERR_CUDA;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { cudaFree(d_sink); return WB_ERR_CUDA; }
  const int grid = wb_sm_count() * 8, threads = 256, iters = 1 << 15;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {   // first repetition = warm-up
    cudaEventRecord(e0, stream);
    WB_LAUNCH("dfma_peak_kernel", dfma_peak_kernel<<<grid, threads, 0, stream>>>(d_sink, iters, 0.999999, 1e-9));
    cudaEventRecord(e1, stream);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d_sink); return WB_ERR_CUDA; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 8.0 * (double)iters * (double)grid * threads / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_sink);
  *tflops_out = best;
  return WB_OK;
}
