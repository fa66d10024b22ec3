// C-ABI entry points (include/worldb200.h).  Thin: argument checks, H2D/D2H staging for
// the host-pointer variants, and calls into the stage runners.
#include "../../include/worldb200.h"

#include <math.h>
#include <string.h>
#include <mutex>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <stdlib.h>
#include <new>
#include <thread>
#include <vector>

#include "wb_internal.h"
#include "wb_harvest.h"

namespace {

std::once_flag g_ctx_once;
int g_ctx_status = WB_OK;
cudaStream_t g_stream = nullptr;

int ctx_init() {
  std::call_once(g_ctx_once, []() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
      fprintf(stderr, "worldb200: no CUDA device available -- this library has no CPU fallback\n");
      g_ctx_status = WB_ERR_CUDA;
      return;
    }
    if (cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking) != cudaSuccess) {
      g_ctx_status = WB_ERR_CUDA;
      return;
    }
    g_ctx_status = wb_rng_init();
  });
  return g_ctx_status;
}

inline cudaStream_t pick_stream(void *stream) { return stream ? (cudaStream_t)stream : g_stream; }

// stand-alone stage calls draw from the process-global randn stream and move it on (reference semantics)
inline WbRngCursor global_cursor() {
  WbRngCursor c;
  c.state = wb_rng_global_state();
  return c;
}

// ---- optional host-side phase trace (WB_TRACE=1): wall-clock milestones of the host-pointer entry points
struct HostTrace {
  bool on;
  const char *what;
  std::chrono::steady_clock::time_point t0, last;
  explicit HostTrace(const char *w) : what(w) {
    static const bool enabled = getenv("WB_TRACE") != nullptr;
    on = enabled;
    if (on) t0 = last = std::chrono::steady_clock::now();
  }
  void mark(const char *phase, cudaStream_t sync_stream = nullptr) {
    if (!on) return;
    if (sync_stream) cudaStreamSynchronize(sync_stream);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[wb_trace] %-22s %-18s +%8.3f ms (total %8.3f)\n", what, phase,
            std::chrono::duration<double, std::milli>(now - last).count(),
            std::chrono::duration<double, std::milli>(now - t0).count());
    last = now;
  }
};

// ---- host <-> device transfers of row-pointer matrices -------------------------------------------
// The reference API hands matrices over as separately allocated rows (test/test.cpp:146-149).  Rows are
// staged through pinned memory in a few chunks: the DMA of chunk k+1 overlaps the (multi-threaded)
// host memcpy of chunk k.
const int kRowChunks = 8;
const int kSynRanges = 4;   // sample ranges of the pipelined host Synthesis::compute
const double kSynRangeEnd[kSynRanges] = {0.30, 0.58, 0.82, 1.0};   // (a range costs ~0.1 ms whatever its size: few, the last one short)

// Persistent helper threads for the row copies (spawning threads per chunk cost more than the copies).
// Workers spin briefly between jobs -- the chunks of one matrix follow each other within microseconds --
// and then sleep on a condition variable.  The calling thread takes a slice too.
class CopyPool {
 public:
  static CopyPool &get() {
    static CopyPool pool;
    return pool;
  }
  // f(begin, end) over contiguous slices of [r0, r1); returns when every slice is done
  void parallel_for(int r0, int r1, const std::function<void(int, int)> &f) {
    const int n = r1 - r0;
    const int parts = std::max(1, std::min((int)workers_.size() + 1, n / 32));
    if (parts == 1) { f(r0, r1); return; }
    std::lock_guard<std::mutex> serial(call_mutex_);   // one job at a time (entry points of different handles may race)
    Job job;
    {
      std::lock_guard<std::mutex> lock(m_);
      job.f = &f; job.r0 = r0; job.n = n; job.parts = parts; job.gen = ++generation_value_;
      job_ = job;
      pending_.store(parts, std::memory_order_relaxed);
      ticket_.store((job.gen << 32) | 0ull, std::memory_order_release);
      generation_.store(job.gen, std::memory_order_release);
    }
    cv_.notify_all();
    run_parts(job);
    while (pending_.load(std::memory_order_acquire) > 0) std::this_thread::yield();
  }

 private:
  struct Job {
    const std::function<void(int, int)> *f = nullptr;
    int r0 = 0, n = 0, parts = 0;
    unsigned long long gen = 0;
  };
  // Helper threads = this process's share of the host's cores (one process per GPU: LOCAL_WORLD_SIZE of them
  // share the box; WB_COPY_THREADS overrides), minus the calling thread, at most 11.  Eight ranks with 11 spinning
  // helpers each on a 32-core box were what collapsed the host API at N = 8 in round 1.
  CopyPool() {
    unsigned hc = std::thread::hardware_concurrency();
    unsigned ranks = 1;
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) { const int v = atoi(e); if (v > 1) ranks = (unsigned)v; }
    unsigned share = hc / ranks;
    int n = (int)std::min(share > 1 ? share - 1 : 0u, 11u);
    if (const char *e = getenv("WB_COPY_THREADS")) { const int v = atoi(e); if (v >= 1 && v <= 64) n = v - 1; }
    for (int i = 0; i < n; ++i) workers_.emplace_back([this]() { worker(); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lock(m_);
      stop_ = true;
      generation_.store(++generation_value_, std::memory_order_release);
    }
    cv_.notify_all();
    for (auto &t : workers_) t.join();
  }
  // Parts are handed out through one (generation, next part) word with compare-and-swap, against a
  // private snapshot of the job: a worker still looping on an old job can neither take nor skip a part of
  // the next one.
  void run_parts(const Job &job) {
    for (;;) {
      unsigned long long cur = ticket_.load(std::memory_order_acquire);
      if ((cur >> 32) != job.gen) return;
      const int part = (int)(cur & 0xffffffffull);
      if (part >= job.parts) return;
      if (!ticket_.compare_exchange_weak(cur, cur + 1, std::memory_order_acq_rel)) continue;
      const int b = job.r0 + (int)((long long)job.n * part / job.parts);
      const int e = job.r0 + (int)((long long)job.n * (part + 1) / job.parts);
      (*job.f)(b, e);
      pending_.fetch_sub(1, std::memory_order_acq_rel);
    }
  }
  void worker() {
    unsigned long long seen = 0;
    for (;;) {
      // a short spin (the chunks of one matrix follow each other within microseconds), then sleep
      bool got = false;
      const auto t0 = std::chrono::steady_clock::now();
      while (std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(40)) {
        if (generation_.load(std::memory_order_acquire) != seen) { got = true; break; }
      }
      Job job;
      {
        std::unique_lock<std::mutex> lock(m_);
        if (!got) cv_.wait(lock, [&]() { return generation_.load(std::memory_order_acquire) != seen; });
        seen = generation_.load(std::memory_order_acquire);
        if (stop_) return;
        job = job_;
      }
      run_parts(job);
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_, call_mutex_;
  std::condition_variable cv_;
  std::atomic<unsigned long long> generation_{0}, ticket_{0};
  unsigned long long generation_value_ = 0;
  std::atomic<int> pending_{0};
  Job job_;
  bool stop_ = false;
};

template <typename F>
void parallel_rows(int r0, int r1, F f) {
  CopyPool::get().parallel_for(r0, r1, [&](int b, int e) { for (int i = b; i < e; ++i) f(i); });
}

// The reference's callers allocate every row separately (test/test.cpp:146-149), but a matrix handed over as
// row pointers into ONE block (numpy / torch / a C++ arena) is just as common.  If the rows are contiguous AND the
// block is page-locked (cudaHostAlloc / cudaHostRegister by its owner, torch pin_memory()), the DMA engine reads /
// writes the caller's memory directly: no bounce buffer, no host memcpy, no helper threads.
template <typename P>
bool rows_contiguous(P rows, int n_rows, int cols) {
  for (int i = 1; i < n_rows; ++i)
    if (rows[i] != rows[i - 1] + cols) return false;
  return n_rows > 0;
}
bool host_range_pinned(const void *ptr, size_t bytes) {
  if (!ptr || bytes == 0) return false;
  cudaPointerAttributes a0, a1;
  if (cudaPointerGetAttributes(&a0, ptr) != cudaSuccess || cudaPointerGetAttributes(&a1, (const char *)ptr + bytes - 1) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a0.type == cudaMemoryTypeHost && a1.type == cudaMemoryTypeHost;
}
template <typename P>
bool rows_direct(P rows, int n_rows, int cols) {
  return rows_contiguous(rows, n_rows, cols) && host_range_pinned(rows[0], sizeof(double) * (size_t)n_rows * cols);
}

// copy a contiguous [rows][cols] device matrix into separately allocated host rows
int rows_to_host(WbWorkspace *ws, const double *d_src, int rows, int cols, double **dst, cudaStream_t st) {
  const size_t bytes = sizeof(double) * (size_t)rows * cols;
  if (rows_direct(dst, rows, cols)) {
    WB_CUDA_CHECK(cudaMemcpyAsync(dst[0], d_src, bytes, cudaMemcpyDeviceToHost, st));
    WB_CUDA_CHECK(cudaStreamSynchronize(st));
    return WB_OK;
  }
  double *stage = (double *)ws->get_pinned("rows_stage", bytes);
  if (!stage) return WB_ERR_CUDA;
  cudaEvent_t ev[kRowChunks];
  int bounds[kRowChunks + 1];
  for (int c = 0; c <= kRowChunks; ++c) bounds[c] = (int)((long long)rows * c / kRowChunks);
  for (int c = 0; c < kRowChunks; ++c) {
    WB_CUDA_CHECK(cudaEventCreateWithFlags(&ev[c], cudaEventDisableTiming));
    const size_t off = (size_t)bounds[c] * cols, cnt = (size_t)(bounds[c + 1] - bounds[c]) * cols;
    if (cnt) WB_CUDA_CHECK(cudaMemcpyAsync(stage + off, d_src + off, sizeof(double) * cnt, cudaMemcpyDeviceToHost, st));
    WB_CUDA_CHECK(cudaEventRecord(ev[c], st));
  }
  int rc = WB_OK;
  for (int c = 0; c < kRowChunks; ++c) {
    if (cudaEventSynchronize(ev[c]) != cudaSuccess) rc = WB_ERR_CUDA;
    if (!rc) parallel_rows(bounds[c], bounds[c + 1], [=](int i) { memcpy(dst[i], stage + (size_t)i * cols, sizeof(double) * cols); });
    cudaEventDestroy(ev[c]);
  }
  return rc;
}

// gather separately allocated host rows into a contiguous device matrix
int rows_to_device(WbWorkspace *ws, const char *name, const double *const *src, int rows, int cols,
                   double *d_dst, cudaStream_t st) {
  const size_t bytes = sizeof(double) * (size_t)rows * cols;
  if (rows_direct(src, rows, cols)) {
    WB_CUDA_CHECK(cudaMemcpyAsync(d_dst, src[0], bytes, cudaMemcpyHostToDevice, st));
    return WB_OK;
  }
  double *stage = (double *)ws->get_pinned(name, bytes);
  if (!stage) return WB_ERR_CUDA;
  for (int c = 0; c < kRowChunks; ++c) {
    const int r0 = (int)((long long)rows * c / kRowChunks), r1 = (int)((long long)rows * (c + 1) / kRowChunks);
    parallel_rows(r0, r1, [=](int i) { memcpy(stage + (size_t)i * cols, src[i], sizeof(double) * cols); });
    const size_t off = (size_t)r0 * cols, cnt = (size_t)(r1 - r0) * cols;
    if (cnt) WB_CUDA_CHECK(cudaMemcpyAsync(d_dst + off, stage + off, sizeof(double) * cnt, cudaMemcpyHostToDevice, st));
  }
  return WB_OK;
}

// Streams / events for a row-chunked compute + download (WbRowChunks): frame ranges alternate between the
// library stream and `alt`, finished ranges are downloaded on `copy` and scattered to the caller's rows.
struct RowPipeline {
  cudaStream_t alt = nullptr, copy = nullptr;
  cudaEvent_t ev[kRowChunks] = {nullptr}, done[kRowChunks] = {nullptr}, ready = nullptr;
  bool ok = false;
  int init() {
    if (ok) return WB_OK;
    if (cudaStreamCreateWithFlags(&alt, cudaStreamNonBlocking) != cudaSuccess) return WB_ERR_CUDA;
    if (cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking) != cudaSuccess) return WB_ERR_CUDA;
    if (cudaEventCreateWithFlags(&ready, cudaEventDisableTiming) != cudaSuccess) return WB_ERR_CUDA;
    for (int c = 0; c < kRowChunks; ++c) {
      if (cudaEventCreateWithFlags(&ev[c], cudaEventDisableTiming) != cudaSuccess) return WB_ERR_CUDA;
      if (cudaEventCreateWithFlags(&done[c], cudaEventDisableTiming) != cudaSuccess) return WB_ERR_CUDA;
    }
    ok = true;
    return WB_OK;
  }
  ~RowPipeline() {
    for (int c = 0; c < kRowChunks; ++c) {
      if (ev[c]) cudaEventDestroy(ev[c]);
      if (done[c]) cudaEventDestroy(done[c]);
    }
    if (ready) cudaEventDestroy(ready);
    if (alt) cudaStreamDestroy(alt);
    if (copy) cudaStreamDestroy(copy);
  }
  void describe(int rows, WbRowChunks *ch) const {
    ch->n = kRowChunks;
    for (int c = 0; c <= kRowChunks; ++c) ch->bounds[c] = (int)((long long)rows * c / kRowChunks);
    for (int c = 0; c < kRowChunks; ++c) ch->ev[c] = ev[c];
    ch->alt = alt;
    ch->ev_ready = ready;
  }
  // download range c as soon as ev[c] fires, scatter it to the caller's rows while the next ranges compute
  int download(WbWorkspace *ws, const WbRowChunks &ch, const double *d_src, int cols, double **dst) {
    const int rows = ch.bounds[ch.n];
    if (rows_direct(dst, rows, cols)) {
      // straight into the caller's page-locked block, range by range as the frame kernels finish
      for (int c = 0; c < ch.n; ++c) {
        const size_t off = (size_t)ch.bounds[c] * cols, cnt = (size_t)(ch.bounds[c + 1] - ch.bounds[c]) * cols;
        WB_CUDA_CHECK(cudaStreamWaitEvent(copy, ch.ev[c], 0));
        if (cnt) WB_CUDA_CHECK(cudaMemcpyAsync(dst[0] + off, d_src + off, sizeof(double) * cnt, cudaMemcpyDeviceToHost, copy));
      }
      WB_CUDA_CHECK(cudaStreamSynchronize(copy));
      return WB_OK;
    }
    double *stage = (double *)ws->get_pinned("rows_stage", sizeof(double) * (size_t)rows * cols);
    if (!stage) return WB_ERR_CUDA;
    for (int c = 0; c < ch.n; ++c) {
      const size_t off = (size_t)ch.bounds[c] * cols, cnt = (size_t)(ch.bounds[c + 1] - ch.bounds[c]) * cols;
      WB_CUDA_CHECK(cudaStreamWaitEvent(copy, ch.ev[c], 0));
      if (cnt) WB_CUDA_CHECK(cudaMemcpyAsync(stage + off, d_src + off, sizeof(double) * cnt, cudaMemcpyDeviceToHost, copy));
      WB_CUDA_CHECK(cudaEventRecord(done[c], copy));
    }
    int rc = WB_OK;
    for (int c = 0; c < ch.n; ++c) {
      if (cudaEventSynchronize(done[c]) != cudaSuccess) rc = WB_ERR_CUDA;
      if (!rc) parallel_rows(ch.bounds[c], ch.bounds[c + 1], [=](int i) { memcpy(dst[i], stage + (size_t)i * cols, sizeof(double) * cols); });
    }
    return rc;
  }
};

int vec_to_device(WbWorkspace *ws, const char *name, const double *src, size_t n, double **d_out, cudaStream_t st) {
  double *d = (double *)ws->get(name, sizeof(double) * n);
  if (!d) return WB_ERR_CUDA;
  WB_CUDA_CHECK(cudaMemcpyAsync(d, src, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  *d_out = d;
  return WB_OK;
}

}  // namespace

struct wb_cheaptrick {
  int fs;
  WbCheapTrickOption opt;   // fft_size resolved
  double f0_floor_internal; // cheaptrick.cpp:34 / :44
  WbWorkspace ws;
  RowPipeline rows;
};

struct wb_harvest {
  WbHarvestPlan plan;
  WbWorkspace ws;
};

struct wb_pipeline {
  int fs;
  WbHarvestPlan plan;
  WbCheapTrickOption ct;
  double ct_f0_floor_internal;
  WbD4COption d4c;
  WbWorkspace ws;
  // After Harvest the chain forks: CheapTrick stays on the caller's stream, D4C runs on `d4c_stream`
  // and the Synthesis time base / pulse list / noise on `side`; they join before the impulse responses.
  cudaStream_t side = nullptr, d4c_stream = nullptr;
  WbStageSplit d4c_split;            // second stream of the D4C stage (halves of the rows, see WbStageSplit)
  unsigned long long graph_generation = 0;   // ws.generation() when the graph was captured
  int stream_f0_length = 0;          // sharded streams: the length given to wb_pipeline_stream_begin_dev
  int stream_samples[2] = {0, 0};    // sharded streams: the sample range the time base of this rank was built for
  double stream_f0_bound = 0.0;      // sharded streams: upper bound of an external f0 contour (<= 0: f0_ceil * 1.25)
  cudaEvent_t ev_f0 = nullptr, ev_tb = nullptr, ev_ct_count = nullptr, ev_body_count = nullptr, ev_d4c = nullptr,
              ev_start = nullptr;
  // CUDA graph of one whole run (captured after a warm run with the same arguments)
  bool use_graph = false;
  cudaGraphExec_t graph_exec = nullptr;
  unsigned long long graph_kernels = 0;
  const void *graph_key[13] = {nullptr}, *warm_key[13] = {nullptr};
  // wb_pipeline_run with page-locked host outputs: every result goes home on `copy` as soon as its stage is done
  // (f0 after Harvest, the spectrogram beside D4C, the aperiodicity beside the impulse responses) instead of after
  // the whole chain; null pointers = no in-chain downloads
  struct HostOut { double *tpos = nullptr, *f0 = nullptr, *sp = nullptr, *ap = nullptr, *y = nullptr; } host_out;
  cudaStream_t copy = nullptr;
  cudaEvent_t ev_ct_done = nullptr, ev_copy_done = nullptr;
  int graph_len[2] = {0, 0};
  int warm_len[2] = {0, 0};
  double mod_f0_shift = NAN, mod_ratio = 0.0;   // parameter modification between analysis and synthesis (off)
  bool private_rng = false;          // batch mode: every run starts from the reference's seed
  WbRngState *d_rng_private = nullptr;
  WbRngState *d_rng_seed = nullptr;
  ~wb_pipeline() {
    if (ev_f0) cudaEventDestroy(ev_f0);
    if (ev_tb) cudaEventDestroy(ev_tb);
    if (ev_ct_count) cudaEventDestroy(ev_ct_count);
    if (ev_body_count) cudaEventDestroy(ev_body_count);
    if (ev_d4c) cudaEventDestroy(ev_d4c);
    if (ev_start) cudaEventDestroy(ev_start);
    if (ev_ct_done) cudaEventDestroy(ev_ct_done);
    if (ev_copy_done) cudaEventDestroy(ev_copy_done);
    if (copy) cudaStreamDestroy(copy);
    if (side) cudaStreamDestroy(side);
    if (d4c_stream) cudaStreamDestroy(d4c_stream);
    if (d4c_split.alt) cudaStreamDestroy(d4c_split.alt);
    for (int k = 0; k < 2; ++k) {
      if (d4c_split.fork[k]) cudaEventDestroy(d4c_split.fork[k]);
      if (d4c_split.join[k]) cudaEventDestroy(d4c_split.join[k]);
    }
    if (d_rng_private) cudaFree(d_rng_private);
    if (d_rng_seed) cudaFree(d_rng_seed);
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
  }
};

struct wb_d4c {
  int fs;
  WbD4COption opt;
  WbWorkspace ws;
  RowPipeline rows;
};

struct wb_synthesis {
  int fs;
  int fft_size;
  double frame_period_ms;
  WbWorkspace ws;
  // host-pointer compute(): the pulse list (f0 only) is built on `side` while sp / ap are still uploading; page-locked
  // contiguous matrices are uploaded on `copy` in row ranges and every sample range is rendered as soon as its rows
  // have landed
  cudaStream_t side = nullptr, copy = nullptr;
  cudaEvent_t ev_f0 = nullptr, ev_tb = nullptr, ev_rows[kSynRanges] = {nullptr}, ev_start = nullptr;
  ~wb_synthesis() {
    if (ev_f0) cudaEventDestroy(ev_f0);
    if (ev_tb) cudaEventDestroy(ev_tb);
    if (ev_start) cudaEventDestroy(ev_start);
    for (int c = 0; c < kSynRanges; ++c) if (ev_rows[c]) cudaEventDestroy(ev_rows[c]);
    if (side) cudaStreamDestroy(side);
    if (copy) cudaStreamDestroy(copy);
  }
};

// option checks shared by wb_harvest_create and wb_pipeline_create (NaNs fail every comparison)
static int check_harvest_option(const WbHarvestOption &o) {
  if (o.use_cos_table) return WB_ERR_UNSUPPORTED;  // approximate window table: not offered (exact path only)
  if (!(o.f0_floor > 0) || !(o.f0_ceil > o.f0_floor) || !(o.frame_period > 0) || !(o.target_fs > 0) ||
      !(o.channels_in_octave > 0) || !(o.f0_ceil < 1e6) || !(o.frame_period < 1e6) || !(o.channels_in_octave < 1e4))
    return WB_ERR_ARG;
  return WB_OK;
}
static bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

extern "C" {

int wb_init(int device) {
  if (device >= 0) {
    if (cudaSetDevice(device) != cudaSuccess) return WB_ERR_CUDA;
  }
  return ctx_init();
}

const char *wb_version(void) { return "worldb200 0.1 (sm_100a)"; }

int wb_device_synchronize(void) {
  int rc = ctx_init();
  if (rc) return rc;
  WB_CUDA_CHECK(cudaStreamSynchronize(g_stream));
  return WB_OK;
}

void wb_harvest_option_default(WbHarvestOption *o) {
  o->f0_floor = 71.0; o->f0_ceil = 800.0; o->frame_period = 5.0;
  o->target_fs = 8000.0; o->channels_in_octave = 40.0; o->use_cos_table = 0;
}
void wb_cheaptrick_option_default(WbCheapTrickOption *o) { o->q1 = -0.15; o->f0_floor = 71.0; o->fft_size = 0; }
void wb_d4c_option_default(WbD4COption *o) { o->threshold = 0.85; }

// ---- randn -----------------------------------------------------------------------------
int wb_randn_reseed(void) {
  int rc = ctx_init();
  if (rc) return rc;
  const unsigned int s0[4] = {123456789u, 362436069u, 521288629u, 88675123u};
  return wb_randn_set_state(s0);
}

int wb_randn_get_state(unsigned int state[4]) {
  int rc = ctx_init();
  if (rc) return rc;
  WB_CUDA_CHECK(cudaStreamSynchronize(g_stream));
  WB_CUDA_CHECK(cudaMemcpy(state, wb_rng_global_state(), sizeof(WbRngState), cudaMemcpyDeviceToHost));
  return WB_OK;
}

int wb_randn_set_state(const unsigned int state[4]) {
  int rc = ctx_init();
  if (rc) return rc;
  WB_CUDA_CHECK(cudaStreamSynchronize(g_stream));
  WB_CUDA_CHECK(cudaMemcpy(wb_rng_global_state(), state, sizeof(WbRngState), cudaMemcpyHostToDevice));
  return WB_OK;
}

int wb_randn_skip(unsigned long long n_calls) {
  unsigned int s[4];
  int rc = wb_randn_get_state(s);
  if (rc) return rc;
  wb_rng_host_jump(s, n_calls);
  return wb_randn_set_state(s);
}

int wb_randn_fill(double *out, int n) {
  int rc = ctx_init();
  if (rc) return rc;
  if (n < 0 || (n > 0 && !out)) return WB_ERR_ARG;
  if (n == 0) return WB_OK;
  double *d = nullptr;
  unsigned long long *d_n = nullptr;
  WB_CUDA_CHECK(cudaMalloc(&d, sizeof(double) * n));
  WB_CUDA_CHECK(cudaMalloc(&d_n, sizeof(unsigned long long)));
  const unsigned long long nn = (unsigned long long)n;
  WB_CUDA_CHECK(cudaMemcpyAsync(d_n, &nn, sizeof(nn), cudaMemcpyHostToDevice, g_stream));
  rc = wb_rng_fill(wb_rng_global_state(), nullptr, nullptr, nn, d, g_stream);
  if (!rc) rc = wb_rng_advance(wb_rng_global_state(), d_n, nullptr, g_stream);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(out, d, sizeof(double) * n, cudaMemcpyDeviceToHost, g_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
    if (e != cudaSuccess) rc = WB_ERR_CUDA;
  }
  cudaFree(d);
  cudaFree(d_n);
  return rc;
}

// ---- stand-alone FFT (world_fft.hpp:33-41) -----------------------------------------------
static int fft_host(int kind, const void *in, size_t in_bytes, int n, int batch, void *out, size_t out_bytes) {
  int rc = ctx_init();
  if (rc) return rc;
  if (!in || !out || n <= 0 || batch < 0) return WB_ERR_ARG;
  if (batch == 0) return WB_OK;
  void *d_in = nullptr, *d_out = nullptr;
  WB_CUDA_CHECK(cudaMalloc(&d_in, in_bytes));
  if (cudaMalloc(&d_out, out_bytes) != cudaSuccess) { cudaFree(d_in); return WB_ERR_CUDA; }
  cudaError_t e = cudaMemcpyAsync(d_in, in, in_bytes, cudaMemcpyHostToDevice, g_stream);
  if (e == cudaSuccess) {
    rc = wb_fft_batch_dev(kind, d_in, n, batch, d_out, g_stream);
    if (!rc) e = cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, g_stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
  cudaFree(d_in);
  cudaFree(d_out);
  if (rc) return rc;
  return e == cudaSuccess ? WB_OK : WB_ERR_CUDA;
}

int wb_fft_r2c(const double *in, int n, int batch, double *out) {
  return fft_host(0, in, sizeof(double) * (size_t)n * batch, n, batch, out, 2 * sizeof(double) * (size_t)(n / 2 + 1) * batch);
}
int wb_fft_c2r(const double *in, int n, int batch, double *out) {
  return fft_host(1, in, 2 * sizeof(double) * (size_t)(n / 2 + 1) * batch, n, batch, out, sizeof(double) * (size_t)n * batch);
}
int wb_fft_c2c(const double *in, int n, int batch, int sign, double *out) {
  if (sign != 1 && sign != 2) return WB_ERR_ARG;
  return fft_host(sign == 1 ? 2 : 3, in, 2 * sizeof(double) * (size_t)n * batch, n, batch, out, 2 * sizeof(double) * (size_t)n * batch);
}

// ---- CheapTrick ------------------------------------------------------------------------
int wb_cheaptrick_get_fft_size(int fs, double f0_floor) {
  // cheaptrick.cpp:97-100, evaluated on the host with the reference's expression
  return (int)pow(2.0, 1.0 + (int)(log(3.0 * fs / f0_floor + 1) / WB_LOG2));
}

double wb_cheaptrick_get_f0_floor(int fs, int fft_size) { return 3 * fs / (fft_size - 3.0); }

int wb_cheaptrick_create(int fs, const WbCheapTrickOption *opt, wb_cheaptrick_t **out) {
  if (!out || fs <= 0) return WB_ERR_ARG;
  int rc = ctx_init();
  if (rc) return rc;
  wb_cheaptrick *h = new (std::nothrow) wb_cheaptrick();
  if (!h) return WB_ERR_ARG;
  h->fs = fs;
  wb_cheaptrick_option_default(&h->opt);
  if (opt) { h->opt.q1 = opt->q1; h->opt.f0_floor = opt->f0_floor; h->opt.fft_size = opt->fft_size; }
  if (!(h->opt.f0_floor > 0) || h->opt.fft_size < 0 ||
      (h->opt.fft_size != 0 && (!is_pow2(h->opt.fft_size) || h->opt.fft_size < 128 || h->opt.fft_size > 16384))) {
    delete h;
    return WB_ERR_ARG;
  }
  if (h->opt.fft_size == 0) h->opt.fft_size = wb_cheaptrick_get_fft_size(fs, h->opt.f0_floor);
  h->f0_floor_internal = wb_cheaptrick_get_f0_floor(fs, h->opt.fft_size);
  *out = h;
  return WB_OK;
}

void wb_cheaptrick_destroy(wb_cheaptrick_t *h) { delete h; }

int wb_cheaptrick_fft_size(const wb_cheaptrick_t *h) { return h ? h->opt.fft_size : 0; }

int wb_cheaptrick_compute_dev(wb_cheaptrick_t *h, const double *d_x, int x_length, const double *d_tpos,
                              const double *d_f0, int f0_length, double *d_sp, void *stream) {
  if (!h || !d_x || !d_tpos || !d_f0 || !d_sp || x_length <= 0 || f0_length < 0) return WB_ERR_ARG;
  return wb_cheaptrick_run(&h->ws, h->fs, h->opt.fft_size, h->opt.q1, h->f0_floor_internal, d_x, x_length,
                           d_tpos, d_f0, f0_length, d_sp, global_cursor(), pick_stream(stream));
}

int wb_cheaptrick_compute(wb_cheaptrick_t *h, const double *x, int x_length, const double *tpos,
                          const double *f0, int f0_length, double **spectrogram) {
  if (!h || !x || !tpos || !f0 || !spectrogram || x_length <= 0 || f0_length < 0) return WB_ERR_ARG;
  if (f0_length == 0) return WB_OK;
  cudaStream_t st = g_stream;
  const int bins = h->opt.fft_size / 2 + 1;
  double *d_x, *d_t, *d_f;
  int rc;
  if ((rc = vec_to_device(&h->ws, "h_x", x, x_length, &d_x, st))) return rc;
  if ((rc = vec_to_device(&h->ws, "h_tpos", tpos, f0_length, &d_t, st))) return rc;
  if ((rc = vec_to_device(&h->ws, "h_f0", f0, f0_length, &d_f, st))) return rc;
  double *d_sp = (double *)h->ws.get("h_sp", sizeof(double) * (size_t)f0_length * bins);
  if (!d_sp) return WB_ERR_CUDA;
  HostTrace tr("cheaptrick_compute");
  if ((rc = h->rows.init())) return rc;
  WbRowChunks ch;
  h->rows.describe(f0_length, &ch);
  if ((rc = wb_cheaptrick_run(&h->ws, h->fs, h->opt.fft_size, h->opt.q1, h->f0_floor_internal, d_x, x_length, d_t, d_f,
                              f0_length, d_sp, global_cursor(), st, &ch)))
    return rc;
  tr.mark("enqueued");
  if ((rc = h->rows.download(&h->ws, ch, d_sp, bins, spectrogram))) return rc;
  tr.mark("rows_to_host");
  return h->ws.read_error_flag(st);
}

// ---- D4C --------------------------------------------------------------------------------
int wb_get_number_of_aperiodicities(int fs) { return wb_number_of_aperiodicities(fs); }

int wb_d4c_create(int fs, const WbD4COption *opt, wb_d4c_t **out) {
  if (!out || fs <= 0) return WB_ERR_ARG;
  int rc = ctx_init();
  if (rc) return rc;
  wb_d4c *h = new (std::nothrow) wb_d4c();
  if (!h) return WB_ERR_ARG;
  h->fs = fs;
  wb_d4c_option_default(&h->opt);
  if (opt) h->opt.threshold = opt->threshold;
  *out = h;
  return WB_OK;
}

void wb_d4c_destroy(wb_d4c_t *h) { delete h; }

int wb_d4c_compute_dev(wb_d4c_t *h, const double *d_x, int x_length, const double *d_tpos, const double *d_f0,
                       int f0_length, int fft_size, double *d_ap, void *stream) {
  if (!h || !d_x || !d_tpos || !d_f0 || !d_ap || x_length <= 0 || f0_length < 0 || fft_size < 2) return WB_ERR_ARG;
  return wb_d4c_run(&h->ws, h->fs, h->opt.threshold, d_x, x_length, d_tpos, d_f0, f0_length, fft_size, d_ap,
                    global_cursor(), pick_stream(stream));
}

int wb_d4c_compute(wb_d4c_t *h, const double *x, int x_length, const double *tpos, const double *f0,
                   int f0_length, int fft_size, double **aperiodicity) {
  if (!h || !x || !tpos || !f0 || !aperiodicity || x_length <= 0 || f0_length < 0 || fft_size < 2) return WB_ERR_ARG;
  if (f0_length == 0) return WB_OK;
  cudaStream_t st = g_stream;
  const int bins = fft_size / 2 + 1;
  double *d_x, *d_t, *d_f;
  int rc;
  if ((rc = vec_to_device(&h->ws, "h_x", x, x_length, &d_x, st))) return rc;
  if ((rc = vec_to_device(&h->ws, "h_tpos", tpos, f0_length, &d_t, st))) return rc;
  if ((rc = vec_to_device(&h->ws, "h_f0", f0, f0_length, &d_f, st))) return rc;
  double *d_ap = (double *)h->ws.get("h_ap", sizeof(double) * (size_t)f0_length * bins);
  if (!d_ap) return WB_ERR_CUDA;
  HostTrace tr("d4c_compute");
  if ((rc = h->rows.init())) return rc;
  WbRowChunks ch;
  h->rows.describe(f0_length, &ch);
  if ((rc = wb_d4c_run(&h->ws, h->fs, h->opt.threshold, d_x, x_length, d_t, d_f, f0_length, fft_size, d_ap,
                       global_cursor(), st, &ch)))
    return rc;
  tr.mark("enqueued");
  if ((rc = h->rows.download(&h->ws, ch, d_ap, bins, aperiodicity))) return rc;
  tr.mark("rows_to_host");
  return h->ws.read_error_flag(st);
}

// ---- Synthesis ---------------------------------------------------------------------------
int wb_synthesis_create(int fs, int fft_size, double frame_period_ms, wb_synthesis_t **out) {
  if (!out || fs <= 0 || fft_size <= 0 || !(frame_period_ms > 0)) return WB_ERR_ARG;
  int rc = ctx_init();
  if (rc) return rc;
  wb_synthesis *h = new (std::nothrow) wb_synthesis();
  if (!h) return WB_ERR_ARG;
  h->fs = fs; h->fft_size = fft_size; h->frame_period_ms = frame_period_ms;
  bool ok = cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&h->copy, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&h->ev_f0, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&h->ev_tb, cudaEventDisableTiming) == cudaSuccess;
  for (int c = 0; ok && c < kSynRanges; ++c) ok = cudaEventCreateWithFlags(&h->ev_rows[c], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    delete h;
    return WB_ERR_CUDA;
  }
  *out = h;
  return WB_OK;
}

void wb_synthesis_destroy(wb_synthesis_t *h) { delete h; }

int wb_synthesis_compute_dev(wb_synthesis_t *h, const double *d_f0, int f0_length, const double *d_sp,
                             const double *d_ap, int out_length, double *d_out, double f0_upper_bound,
                             void *stream) {
  if (!h || !d_f0 || !d_sp || !d_ap || !d_out || f0_length < 2 || out_length < 0) return WB_ERR_ARG;
  return wb_synthesis_run(&h->ws, h->fs, h->fft_size, h->frame_period_ms, d_f0, f0_length, d_sp, d_ap,
                          out_length, d_out, f0_upper_bound, global_cursor(), pick_stream(stream));
}

int wb_synthesis_compute(wb_synthesis_t *h, const double *f0, int f0_length, const double *const *spectrogram,
                         const double *const *aperiodicity, int out_length, double *out) {
  if (!h || !f0 || !spectrogram || !aperiodicity || !out || f0_length < 2 || out_length < 0) return WB_ERR_ARG;
  if (out_length == 0) return WB_OK;
  cudaStream_t st = g_stream;
  const int bins = h->fft_size / 2 + 1;
  double max_f0 = 0.0;
  for (int i = 0; i < f0_length; ++i) if (f0[i] > max_f0) max_f0 = f0[i];
  double *d_f;
  int rc;
  if ((rc = vec_to_device(&h->ws, "h_f0", f0, f0_length, &d_f, st))) return rc;
  double *d_sp = (double *)h->ws.get("h_sp", sizeof(double) * (size_t)f0_length * bins);
  double *d_ap = (double *)h->ws.get("h_ap", sizeof(double) * (size_t)f0_length * bins);
  double *d_out = (double *)h->ws.get("h_out", sizeof(double) * (size_t)out_length);
  if (!d_sp || !d_ap || !d_out) return WB_ERR_CUDA;
  HostTrace tr("synthesis_compute");
  const double hop = h->frame_period_ms / 1000.0 * h->fs;   // samples per frame
  if (rows_direct(spectrogram, f0_length, bins) && rows_direct(aperiodicity, f0_length, bins) && hop >= 1.0 &&
      out_length >= 8 * kSynRanges * h->fft_size) {
    // Page-locked contiguous matrices: the rows go up on the copy stream in kSynRanges row ranges (spectrogram and
    // aperiodicity of a range back to back) and every sample range is rendered -- its own pulses at their
    // whole-waveform noise positions, like a shard of a long stream -- as soon as the rows its pulses interpolate
    // between have landed: the impulse responses run under the remaining uploads.
    WbRngCursor cur = global_cursor();
    cur.advance = false;
    // (WB_TRACE: device-side timeline of the call)
    cudaEvent_t tev[4 + 2 * kSynRanges] = {nullptr};
    if (tr.on) for (auto &e : tev) cudaEventCreate(&e);
    if (tr.on) cudaEventRecord(tev[0], st);
    if ((rc = wb_synthesis_prepare(&h->ws, h->fft_size, st))) return rc;
    WB_CUDA_CHECK(cudaEventRecord(h->ev_f0, st));
    WB_CUDA_CHECK(cudaStreamWaitEvent(h->side, h->ev_f0, 0));
    if ((rc = wb_synthesis_timebase(&h->ws, h->fs, h->fft_size, h->frame_period_ms, d_f, f0_length, out_length, h->side, nullptr)))
      return rc;
    WB_CUDA_CHECK(cudaEventRecord(h->ev_tb, h->side));
    if (tr.on) cudaEventRecord(tev[1], h->side);
    WB_CUDA_CHECK(cudaEventRecord(h->ev_start, st));            // (d_sp / d_ap may still be read by the previous call's kernels)
    WB_CUDA_CHECK(cudaStreamWaitEvent(h->copy, h->ev_start, 0));
    int row_done = 0;
    int s_end[kSynRanges];
    for (int c = 0; c < kSynRanges; ++c) {
      s_end[c] = c == kSynRanges - 1 ? out_length : (int)(out_length * kSynRangeEnd[c]);
      // every frame a pulse reaching into [.., s_end) interpolates between (see StreamPlan.rows)
      int row_hi = (c == kSynRanges - 1) ? f0_length : (int)((s_end[c] + h->fft_size) / hop) + 3;
      if (row_hi > f0_length) row_hi = f0_length;
      if (row_hi > row_done) {
        const size_t off = (size_t)row_done * bins, cnt = (size_t)(row_hi - row_done) * bins;
        WB_CUDA_CHECK(cudaMemcpyAsync(d_sp + off, spectrogram[0] + off, sizeof(double) * cnt, cudaMemcpyHostToDevice, h->copy));
        WB_CUDA_CHECK(cudaMemcpyAsync(d_ap + off, aperiodicity[0] + off, sizeof(double) * cnt, cudaMemcpyHostToDevice, h->copy));
        row_done = row_hi;
      }
      WB_CUDA_CHECK(cudaEventRecord(h->ev_rows[c], h->copy));
      if (tr.on) cudaEventRecord(tev[4 + c], h->copy);
    }
    tr.mark("uploads enqueued");
    // The ranges alternate between the library stream and the (by then idle) time-base stream, a scratch set each:
    // a range is rendered while the one before it still is, and goes home as soon as it is complete, so after the
    // last rows have landed only one short range and its download remain.
    WB_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_tb, 0));
    for (int c = 0; c < kSynRanges; ++c) {
      const int s_begin = c == 0 ? 0 : s_end[c - 1];
      cudaStream_t cs = (c & 1) ? h->side : st;
      if ((rc = wb_synthesis_render_range(&h->ws, h->fs, h->fft_size, h->frame_period_ms, f0_length, d_sp, d_ap, 0, f0_length,
                                          out_length, s_begin, s_end[c], d_out + s_begin, max_f0 + 1.0, cur, cs, c & 1, h->ev_rows[c])))
        return rc;
      if (s_end[c] > s_begin)
        WB_CUDA_CHECK(cudaMemcpyAsync(out + s_begin, d_out + s_begin, sizeof(double) * (size_t)(s_end[c] - s_begin), cudaMemcpyDeviceToHost, cs));
      if (tr.on) cudaEventRecord(tev[4 + kSynRanges + c], cs);
    }
    WB_CUDA_CHECK(cudaEventRecord(h->ev_f0, h->side));
    WB_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_f0, 0));
    unsigned long long *d_ncount = (unsigned long long *)h->ws.find("syn_ncount");
    if (!d_ncount) return WB_ERR_CUDA;
    if ((rc = wb_rng_advance(cur.state, d_ncount, nullptr, st))) return rc;   // what one compute() call draws
    if (tr.on) cudaEventRecord(tev[2], st);
    WB_CUDA_CHECK(cudaStreamSynchronize(st));
    tr.mark("done");
    if (tr.on) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, tev[0], tev[1]); fprintf(stderr, "[wb_trace]   device: time base done  %7.3f ms\n", ms);
      for (int c = 0; c < kSynRanges; ++c) {
        cudaEventElapsedTime(&ms, tev[0], tev[4 + c]); fprintf(stderr, "[wb_trace]   device: rows %d landed    %7.3f ms\n", c, ms);
        cudaEventElapsedTime(&ms, tev[0], tev[4 + kSynRanges + c]); fprintf(stderr, "[wb_trace]   device: range %d home     %7.3f ms\n", c, ms);
      }
      cudaEventElapsedTime(&ms, tev[0], tev[2]); fprintf(stderr, "[wb_trace]   device: all joined      %7.3f ms\n", ms);
      for (auto &e : tev) if (e) cudaEventDestroy(e);
    }
    return h->ws.read_error_flag(st);
  }
  // pulse list + excitation noise on the side stream (needs f0 only), overlapping the row uploads below
  WbRngCursor cur = global_cursor();
  WB_CUDA_CHECK(cudaEventRecord(h->ev_f0, st));
  WB_CUDA_CHECK(cudaStreamWaitEvent(h->side, h->ev_f0, 0));
  if ((rc = wb_synthesis_timebase(&h->ws, h->fs, h->fft_size, h->frame_period_ms, d_f, f0_length, out_length, h->side, &cur)))
    return rc;
  WB_CUDA_CHECK(cudaEventRecord(h->ev_tb, h->side));
  if ((rc = rows_to_device(&h->ws, "rows_stage_sp", spectrogram, f0_length, bins, d_sp, st))) return rc;
  if ((rc = rows_to_device(&h->ws, "rows_stage_ap", aperiodicity, f0_length, bins, d_ap, st))) return rc;
  tr.mark("rows staged");
  WB_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_tb, 0));
  if ((rc = wb_synthesis_render(&h->ws, h->fs, h->fft_size, h->frame_period_ms, f0_length, d_sp, d_ap, out_length, d_out,
                                max_f0 + 1.0, cur, st, true)))
    return rc;
  WB_CUDA_CHECK(cudaMemcpyAsync(out, d_out, sizeof(double) * out_length, cudaMemcpyDeviceToHost, st));
  WB_CUDA_CHECK(cudaStreamSynchronize(st));
  tr.mark("done");
  return h->ws.read_error_flag(st);
}

// ---- Harvest -----------------------------------------------------------------------------
int wb_harvest_get_samples(int fs, int x_length, double frame_period) {
  return static_cast<int>(1000.0 * x_length / fs / frame_period) + 1;  // harvest.cpp:173-176
}

int wb_harvest_create(int fs, const WbHarvestOption *opt, wb_harvest_t **out) {
  if (!out || fs <= 0) return WB_ERR_ARG;
  int rc = ctx_init();
  if (rc) return rc;
  WbHarvestOption o;
  wb_harvest_option_default(&o);
  if (opt) o = *opt;
  if ((rc = check_harvest_option(o))) return rc;
  wb_harvest *h = new (std::nothrow) wb_harvest();
  if (!h) return WB_ERR_ARG;
  WbHarvestOptionInternal oi = {o.f0_floor, o.f0_ceil, o.frame_period, o.target_fs, o.channels_in_octave};
  rc = wb_harvest_plan_init(&h->plan, fs, oi);
  if (rc) { delete h; return rc; }
  *out = h;
  return WB_OK;
}

void wb_harvest_destroy(wb_harvest_t *h) { delete h; }

int wb_harvest_compute_dev(wb_harvest_t *h, const double *d_x, int x_length, double *d_tpos, double *d_f0,
                           void *stream) {
  if (!h || !d_x || !d_tpos || !d_f0 || x_length <= 0) return WB_ERR_ARG;
  cudaStream_t st = pick_stream(stream);
  const double fp = h->plan.opt.frame_period;
  int rc, Lb = 0;
  if (fp == 1.0) {  // harvest.cpp:185-189
    if ((rc = wb_harvest_run_basic(&h->plan, &h->ws, d_x, x_length, 1, d_f0, &Lb, st))) return rc;
    return wb_harvest_pick(d_f0, Lb, 1.0, Lb, d_tpos, d_f0, st);  // identity pick; fills temporal positions
  }
  const int Lb_expected = wb_harvest_get_samples(h->plan.fs, x_length, 1.0);
  double *d_basic = (double *)h->ws.get("hv_basic_f0", sizeof(double) * Lb_expected);
  if (!d_basic) return WB_ERR_CUDA;
  if ((rc = wb_harvest_run_basic(&h->plan, &h->ws, d_x, x_length, 1, d_basic, &Lb, st))) return rc;
  const int f0_length = wb_harvest_get_samples(h->plan.fs, x_length, fp);
  return wb_harvest_pick(d_basic, Lb, fp, f0_length, d_tpos, d_f0, st);
}

int wb_harvest_compute(wb_harvest_t *h, const double *x, int x_length, double *tpos, double *f0) {
  if (!h || !x || !tpos || !f0 || x_length <= 0) return WB_ERR_ARG;
  cudaStream_t st = g_stream;
  const int f0_length = wb_harvest_get_samples(h->plan.fs, x_length, h->plan.opt.frame_period);
  double *d_x;
  int rc;
  if ((rc = vec_to_device(&h->ws, "h_x", x, x_length, &d_x, st))) return rc;
  double *d_t = (double *)h->ws.get("h_tpos", sizeof(double) * f0_length);
  double *d_f = (double *)h->ws.get("h_f0", sizeof(double) * f0_length);
  if (!d_t || !d_f) return WB_ERR_CUDA;
  HostTrace tr("harvest_compute");
  if ((rc = wb_harvest_compute_dev(h, d_x, x_length, d_t, d_f, st))) return rc;
  tr.mark("enqueued");
  tr.mark("kernels", st);
  WB_CUDA_CHECK(cudaMemcpyAsync(tpos, d_t, sizeof(double) * f0_length, cudaMemcpyDeviceToHost, st));
  WB_CUDA_CHECK(cudaMemcpyAsync(f0, d_f, sizeof(double) * f0_length, cudaMemcpyDeviceToHost, st));
  WB_CUDA_CHECK(cudaStreamSynchronize(st));
  return h->ws.read_error_flag(st);
}

/* debug: copy the first n_bytes of a named internal device buffer of the last compute() */
int wb_harvest_debug_read(wb_harvest_t *h, const char *name, void *out, unsigned long long n_bytes) {
  if (!h || !name || !out) return WB_ERR_ARG;
  size_t have = 0;
  void *d = h->ws.find(name, &have);
  if (!d || n_bytes > have) return WB_ERR_ARG;   // unknown buffer, or more than it holds
  WB_CUDA_CHECK(cudaStreamSynchronize(g_stream));
  WB_CUDA_CHECK(cudaMemcpy(out, d, n_bytes, cudaMemcpyDeviceToHost));
  return WB_OK;
}

// ---- whole chain, device resident (test/test.cpp:288-384 call sequence) -----------------------
int wb_pipeline_create(int fs, const WbHarvestOption *hopt, const WbCheapTrickOption *copt, const WbD4COption *dopt,
                       wb_pipeline_t **out) {
  if (!out || fs <= 0) return WB_ERR_ARG;
  int rc = ctx_init();
  if (rc) return rc;
  WbHarvestOption ho;
  wb_harvest_option_default(&ho);
  if (hopt) ho = *hopt;
  if ((rc = check_harvest_option(ho))) return rc;
  if (copt && (!(copt->f0_floor > 0) || copt->fft_size < 0 || (copt->fft_size != 0 && (!is_pow2(copt->fft_size) || copt->fft_size < 128 || copt->fft_size > 16384))))
    return WB_ERR_ARG;
  if (dopt && !(dopt->threshold >= 0.0 && dopt->threshold <= 1.0)) return WB_ERR_ARG;
  wb_pipeline *p = new (std::nothrow) wb_pipeline();
  if (!p) return WB_ERR_ARG;
  p->fs = fs;
  WbHarvestOptionInternal oi = {ho.f0_floor, ho.f0_ceil, ho.frame_period, ho.target_fs, ho.channels_in_octave};
  rc = wb_harvest_plan_init(&p->plan, fs, oi);
  if (rc) { delete p; return rc; }
  wb_cheaptrick_option_default(&p->ct);
  if (copt) p->ct = *copt;
  if (p->ct.fft_size == 0) p->ct.fft_size = wb_cheaptrick_get_fft_size(fs, p->ct.f0_floor);
  p->ct_f0_floor_internal = wb_cheaptrick_get_f0_floor(fs, p->ct.fft_size);
  wb_d4c_option_default(&p->d4c);
  if (dopt) p->d4c = *dopt;
  // D4C is the longest branch after the fork: give its stream the highest priority
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithPriority(&p->d4c_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
      cudaStreamCreateWithPriority(&p->d4c_split.alt, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->d4c_split.fork[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->d4c_split.fork[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->d4c_split.join[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->d4c_split.join[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->ev_f0, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->ev_ct_count, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->ev_body_count, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->ev_d4c, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithFlags(&p->copy, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->ev_ct_done, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->ev_copy_done, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->ev_tb, cudaEventDisableTiming) != cudaSuccess) {
    delete p;
    return WB_ERR_CUDA;
  }
  *out = p;
  return WB_OK;
}

void wb_pipeline_destroy(wb_pipeline_t *p) { delete p; }

/* Batch mode: with fresh != 0 every run of this pipeline draws from its own randn() stream that
 * restarts at the reference's seed, i.e. each utterance is processed like one reference process.
 * Pipelines in this mode are independent of each other and may run concurrently on different
 * streams.  Default (0): the process-global stream, like consecutive calls in one reference process. */
int wb_pipeline_set_fresh_rng(wb_pipeline_t *p, int fresh) {
  if (!p) return WB_ERR_ARG;
  if (fresh && !p->d_rng_private) {
    const WbRngState s0 = {{123456789u, 362436069u, 521288629u, 88675123u}};
    WB_CUDA_CHECK(cudaMalloc(&p->d_rng_private, sizeof(WbRngState)));
    WB_CUDA_CHECK(cudaMalloc(&p->d_rng_seed, sizeof(WbRngState)));
    WB_CUDA_CHECK(cudaMemcpy(p->d_rng_seed, &s0, sizeof(s0), cudaMemcpyHostToDevice));
  }
  p->private_rng = fresh != 0;
  if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }   // the captured chain names the state
  return WB_OK;
}
int wb_pipeline_fft_size(const wb_pipeline_t *p) { return p ? p->ct.fft_size : 0; }
int wb_pipeline_f0_length(const wb_pipeline_t *p, int x_length) {
  return p ? wb_harvest_get_samples(p->fs, x_length, p->plan.opt.frame_period) : 0;
}
int wb_pipeline_out_length(const wb_pipeline_t *p, int x_length) {  // test/test.cpp:362-363
  if (!p) return 0;
  const int f0_length = wb_pipeline_f0_length(p, x_length);
  return static_cast<int>((f0_length - 1) * p->plan.opt.frame_period / 1000.0 * p->fs) + 1;
}

static int pipeline_enqueue(wb_pipeline_t *p, const double *d_x, int x_length, double *d_tpos, double *d_f0,
                            double *d_sp, double *d_ap, double *d_y, int y_length, cudaStream_t st);

/* use_graph != 0: after one ordinary run with a given set of arguments, the next run with the same
 * arguments is captured into a CUDA graph and later runs replay it (one launch instead of ~45). */
int wb_pipeline_set_modification(wb_pipeline_t *p, double f0_shift, double ratio) {
  if (!p) return WB_ERR_ARG;
  if (f0_shift == f0_shift && !(f0_shift > 0.0)) return WB_ERR_ARG;
  p->mod_f0_shift = f0_shift;
  p->mod_ratio = ratio > 0.0 ? ratio : 0.0;
  if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }   // the captured chain changes
  return WB_OK;
}

int wb_pipeline_set_graph(wb_pipeline_t *p, int use_graph) {
  if (!p) return WB_ERR_ARG;
  p->use_graph = use_graph != 0;
  return WB_OK;
}

int wb_pipeline_run_dev(wb_pipeline_t *p, const double *d_x, int x_length, double *d_tpos, double *d_f0,
                        double *d_sp, double *d_ap, double *d_y, int y_length, void *stream) {
  if (!p || !d_x || x_length <= 0 || y_length < 0) return WB_ERR_ARG;
  cudaStream_t st = pick_stream(stream);
  if (!p->use_graph || wb_prof_is_enabled())
    return pipeline_enqueue(p, d_x, x_length, d_tpos, d_f0, d_sp, d_ap, d_y, y_length, st);
  const void *key[13] = {d_x, d_tpos, d_f0, d_sp, d_ap, d_y, (const void *)st, nullptr,
                         p->host_out.tpos, p->host_out.f0, p->host_out.sp, p->host_out.ap, p->host_out.y};
  // (the graph holds raw pointers into the workspace: stale once any of its buffers has been reallocated)
  const bool same = p->graph_exec && memcmp(key, p->graph_key, sizeof(key)) == 0 && p->graph_len[0] == x_length &&
                    p->graph_len[1] == y_length && p->graph_generation == p->ws.generation();
  if (same) {
    WB_CUDA_CHECK(cudaGraphLaunch(p->graph_exec, st));
    wb_launch_counter_add(p->graph_kernels);
    return WB_OK;
  }
  if (p->warm_len[0] != x_length || p->warm_len[1] != y_length || memcmp(key, p->warm_key, sizeof(key)) != 0) {
    // first run with these sizes and buffers: ordinary enqueue (allocations, plan-time tables, lazy initialisation).
    // A graph is captured only when a call repeats the previous one's arguments, so callers that rotate their
    // buffers run on plain launches instead of re-capturing every time.
    int rc = pipeline_enqueue(p, d_x, x_length, d_tpos, d_f0, d_sp, d_ap, d_y, y_length, st);
    if (!rc) { p->warm_len[0] = x_length; p->warm_len[1] = y_length; memcpy(p->warm_key, key, sizeof(key)); }
    return rc;
  }
  if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
  const unsigned long long k0 = wb_launch_counter();
  WB_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  int rc = pipeline_enqueue(p, d_x, x_length, d_tpos, d_f0, d_sp, d_ap, d_y, y_length, st);
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(st, &graph);
  if (rc || e != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    if (rc) return rc;
    // capture not possible: fall back to ordinary launches from now on
    p->use_graph = false;
    return pipeline_enqueue(p, d_x, x_length, d_tpos, d_f0, d_sp, d_ap, d_y, y_length, st);
  }
  p->graph_kernels = wb_launch_counter() - k0;
  e = cudaGraphInstantiate(&p->graph_exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { p->graph_exec = nullptr; p->use_graph = false; cudaGetLastError(); return pipeline_enqueue(p, d_x, x_length, d_tpos, d_f0, d_sp, d_ap, d_y, y_length, st); }
  memcpy(p->graph_key, key, sizeof(key));
  p->graph_len[0] = x_length; p->graph_len[1] = y_length;
  p->graph_generation = p->ws.generation();
  WB_CUDA_CHECK(cudaGraphLaunch(p->graph_exec, st));
  return WB_OK;
}

static int pipeline_enqueue(wb_pipeline_t *p, const double *d_x, int x_length, double *d_tpos, double *d_f0,
                            double *d_sp, double *d_ap, double *d_y, int y_length, cudaStream_t st) {
  const int fs = p->fs;
  const double fp = p->plan.opt.frame_period;
  const int f0_length = wb_pipeline_f0_length(p, x_length);
  const int bins = p->ct.fft_size / 2 + 1;
  if (!d_tpos) d_tpos = (double *)p->ws.get("pl_tpos", sizeof(double) * f0_length);
  if (!d_f0) d_f0 = (double *)p->ws.get("pl_f0", sizeof(double) * f0_length);
  if (!d_sp) d_sp = (double *)p->ws.get("pl_sp", sizeof(double) * (size_t)f0_length * bins);
  if (!d_ap) d_ap = (double *)p->ws.get("pl_ap", sizeof(double) * (size_t)f0_length * bins);
  if (!d_y && y_length > 0) d_y = (double *)p->ws.get("pl_y", sizeof(double) * (size_t)y_length);
  if (!d_tpos || !d_f0 || !d_sp || !d_ap) return WB_ERR_CUDA;
  int rc, Lb = 0;
  WbRngState *rng = wb_rng_global_state();
  if (p->private_rng) {
    rng = p->d_rng_private;
    WB_CUDA_CHECK(cudaMemcpyAsync(rng, p->d_rng_seed, sizeof(WbRngState), cudaMemcpyDeviceToDevice, st));
  }
  WB_CUDA_CHECK(cudaEventRecord(p->ev_start, st));
  // (side stream, joined with the pulse list further down: warm L2 with the randn jump tables while Harvest runs)
  if (y_length > 0 && !wb_prof_is_enabled()) {
    WB_CUDA_CHECK(cudaStreamWaitEvent(p->side, p->ev_start, 0));
    if ((rc = wb_rng_prefetch_tables(p->side))) return rc;
  }
  // Harvest (always analysed on the 1 ms grid, harvest.cpp:185-204)
  if (fp == 1.0) {
    if ((rc = wb_harvest_run_basic(&p->plan, &p->ws, d_x, x_length, 1, d_f0, &Lb, st))) return rc;
    if ((rc = wb_harvest_pick(d_f0, Lb, 1.0, Lb, d_tpos, d_f0, st))) return rc;
  } else {
    const int Lb_expected = wb_harvest_get_samples(fs, x_length, 1.0);
    double *d_basic = (double *)p->ws.get("hv_basic_f0", sizeof(double) * Lb_expected);
    if (!d_basic) return WB_ERR_CUDA;
    if ((rc = wb_harvest_run_basic(&p->plan, &p->ws, d_x, x_length, 1, d_basic, &Lb, st))) return rc;
    if ((rc = wb_harvest_pick(d_basic, Lb, fp, f0_length, d_tpos, d_f0, st))) return rc;
  }
  // parameter modification (test/test.cpp:318-332 applies it between analysis and synthesis): the analysis
  // stages keep the estimated f0, Synthesis gets the scaled one
  const bool mod_f0 = p->mod_f0_shift == p->mod_f0_shift;
  const double *d_f0_syn = d_f0;
  if (mod_f0) {
    double *d_mod = (double *)p->ws.get("pl_f0_mod", sizeof(double) * f0_length);
    if (!d_mod) return WB_ERR_CUDA;
    if ((rc = wb_parameter_modification_run(d_f0, d_mod, f0_length, nullptr, fs, p->ct.fft_size, p->mod_f0_shift, 0.0, st))) return rc;
    d_f0_syn = d_mod;
  }
  // Fork.  Everything below depends on f0 only until the impulse responses need sp, ap and the pulses.
  // The randn() stream is consumed in the reference's serial order CheapTrick -> Love Train -> D4C body
  // -> Synthesis; each stage counts its draws up front, so the stages run concurrently and chain their
  // stream positions through rng_pos[] (WbRngCursor): [0] after CheapTrick, [1] after D4C.
  // (Per-kernel event timing needs one stream: profiling runs the branches back to back.)
  const bool fork = !wb_prof_is_enabled();
  cudaStream_t s_d4c = fork ? p->d4c_stream : st, s_side = fork ? p->side : st;
  const wb_pipeline::HostOut &ho = p->host_out;
  const bool downloads = ho.tpos || ho.f0 || ho.sp || ho.ap || ho.y;
  cudaStream_t s_copy = fork ? p->copy : st;
  unsigned long long *rng_pos = (unsigned long long *)p->ws.get("pl_rng_pos", sizeof(unsigned long long) * 4);
  if (!rng_pos) return WB_ERR_CUDA;
  WB_CUDA_CHECK(cudaEventRecord(p->ev_f0, st));
  if (fork) {
    WB_CUDA_CHECK(cudaStreamWaitEvent(s_d4c, p->ev_f0, 0));
    if (y_length > 0) WB_CUDA_CHECK(cudaStreamWaitEvent(s_side, p->ev_f0, 0));
    if (downloads) WB_CUDA_CHECK(cudaStreamWaitEvent(s_copy, p->ev_f0, 0));
  }
  if (ho.tpos) WB_CUDA_CHECK(cudaMemcpyAsync(ho.tpos, d_tpos, sizeof(double) * f0_length, cudaMemcpyDeviceToHost, s_copy));
  if (ho.f0 && !mod_f0) WB_CUDA_CHECK(cudaMemcpyAsync(ho.f0, d_f0, sizeof(double) * f0_length, cudaMemcpyDeviceToHost, s_copy));
  // CheapTrick (caller's stream)
  WbRngCursor c_ct;
  c_ct.state = rng; c_ct.skip_out = rng_pos + 0; c_ct.advance = false; c_ct.record_skip_out = p->ev_ct_count;
  if ((rc = wb_cheaptrick_run(&p->ws, fs, p->ct.fft_size, p->ct.q1, p->ct_f0_floor_internal, d_x, x_length, d_tpos,
                              d_f0, f0_length, d_sp, c_ct, st)))
    return rc;
  if (p->mod_ratio > 0.0 &&
      (rc = wb_parameter_modification_run(nullptr, nullptr, f0_length, d_sp, fs, p->ct.fft_size, NAN, p->mod_ratio, st)))
    return rc;
  if (ho.sp) {
    if (fork) {
      WB_CUDA_CHECK(cudaEventRecord(p->ev_ct_done, st));
      WB_CUDA_CHECK(cudaStreamWaitEvent(s_copy, p->ev_ct_done, 0));
    }
    WB_CUDA_CHECK(cudaMemcpyAsync(ho.sp, d_sp, sizeof(double) * (size_t)f0_length * bins, cudaMemcpyDeviceToHost, s_copy));
  }
  // D4C
  WbRngCursor c_d4c;
  c_d4c.state = rng; c_d4c.skip_in = rng_pos + 0; c_d4c.skip_out = rng_pos + 1; c_d4c.advance = false;
  c_d4c.wait_skip_in = p->ev_ct_count; c_d4c.record_skip_out = p->ev_body_count;
  if ((rc = wb_d4c_run(&p->ws, fs, p->d4c.threshold, d_x, x_length, d_tpos, d_f0, f0_length, p->ct.fft_size, d_ap,
                       c_d4c, s_d4c, nullptr, nullptr, 0, nullptr, fork ? &p->d4c_split : nullptr)))
    return rc;
  WB_CUDA_CHECK(cudaEventRecord(p->ev_d4c, s_d4c));
  if (ho.ap) {
    if (fork) WB_CUDA_CHECK(cudaStreamWaitEvent(s_copy, p->ev_d4c, 0));
    WB_CUDA_CHECK(cudaMemcpyAsync(ho.ap, d_ap, sizeof(double) * (size_t)f0_length * bins, cudaMemcpyDeviceToHost, s_copy));
  }
  if (y_length > 0) {
    if (!d_y) return WB_ERR_CUDA;
    // Synthesis, part 1 (pulse list) on the side stream
    WbRngCursor c_syn;
    c_syn.state = rng; c_syn.skip_in = rng_pos + 1; c_syn.wait_skip_in = p->ev_body_count;
    c_syn.advance = true;  // moves the state past the whole chain
    if ((rc = wb_synthesis_timebase(&p->ws, fs, p->ct.fft_size, fp, d_f0_syn, f0_length, y_length, s_side, &c_syn))) return rc;
    WB_CUDA_CHECK(cudaEventRecord(p->ev_tb, s_side));
    // join, then part 2.  Harvest's contour is bounded by f0_ceil up to the smoothing overshoot.
    const double f0_bound = p->plan.opt.f0_ceil * 1.25 * (mod_f0 && p->mod_f0_shift > 1.0 ? p->mod_f0_shift : 1.0);
    WB_CUDA_CHECK(cudaStreamWaitEvent(st, p->ev_tb, 0));
    WB_CUDA_CHECK(cudaStreamWaitEvent(st, p->ev_d4c, 0));
    if ((rc = wb_synthesis_render(&p->ws, fs, p->ct.fft_size, fp, f0_length, d_sp, d_ap, y_length, d_y, f0_bound,
                                  c_syn, st, true)))
      return rc;
  } else {
    WB_CUDA_CHECK(cudaStreamWaitEvent(st, p->ev_d4c, 0));
    if ((rc = wb_rng_advance(rng, rng_pos + 1, nullptr, st))) return rc;
  }
  // the caller sees the modified f0, like the demo's world_parameters after ParameterModification (D4C, the
  // last reader of the estimated contour, has been joined above)
  if (mod_f0) WB_CUDA_CHECK(cudaMemcpyAsync(d_f0, d_f0_syn, sizeof(double) * f0_length, cudaMemcpyDeviceToDevice, st));
  if (downloads) {
    if (ho.y && y_length > 0) WB_CUDA_CHECK(cudaMemcpyAsync(ho.y, d_y, sizeof(double) * (size_t)y_length, cudaMemcpyDeviceToHost, st));
    if (ho.f0 && mod_f0) WB_CUDA_CHECK(cudaMemcpyAsync(ho.f0, d_f0, sizeof(double) * f0_length, cudaMemcpyDeviceToHost, st));
    if (fork) {   // the copy stream rejoins the caller's
      WB_CUDA_CHECK(cudaEventRecord(p->ev_copy_done, s_copy));
      WB_CUDA_CHECK(cudaStreamWaitEvent(st, p->ev_copy_done, 0));
    }
  }
  return WB_OK;
}

int wb_pipeline_run(wb_pipeline_t *p, const double *x, int x_length, double *tpos, double *f0, double *sp,
                    double *ap, double *y, int y_length) {
  if (!p || !x || x_length <= 0 || y_length < 0 || (y_length > 0 && !y)) return WB_ERR_ARG;
  cudaStream_t st = g_stream;
  const int f0_length = wb_pipeline_f0_length(p, x_length);
  const int bins = p->ct.fft_size / 2 + 1;
  double *d_x;
  int rc;
  if ((rc = vec_to_device(&p->ws, "h_x", x, x_length, &d_x, st))) return rc;
  double *d_t = (double *)p->ws.get("pl_tpos", sizeof(double) * f0_length);
  double *d_f = (double *)p->ws.get("pl_f0", sizeof(double) * f0_length);
  double *d_sp = (double *)p->ws.get("pl_sp", sizeof(double) * (size_t)f0_length * bins);
  double *d_ap = (double *)p->ws.get("pl_ap", sizeof(double) * (size_t)f0_length * bins);
  double *d_y = (double *)p->ws.get("pl_y", sizeof(double) * (size_t)(y_length > 0 ? y_length : 1));
  if (!d_t || !d_f || !d_sp || !d_ap || !d_y) return WB_ERR_CUDA;
  // Page-locked outputs are downloaded inside the chain, each as soon as its stage is done (see HostOut)
  const size_t mat = sizeof(double) * (size_t)f0_length * bins;
  const bool in_chain = (!tpos || host_range_pinned(tpos, sizeof(double) * f0_length)) &&
                        (!f0 || host_range_pinned(f0, sizeof(double) * f0_length)) && (!sp || host_range_pinned(sp, mat)) &&
                        (!ap || host_range_pinned(ap, mat)) && (y_length == 0 || host_range_pinned(y, sizeof(double) * (size_t)y_length));
  if (in_chain) {
    p->host_out.tpos = tpos; p->host_out.f0 = f0; p->host_out.sp = sp; p->host_out.ap = ap; p->host_out.y = y_length > 0 ? y : nullptr;
    rc = wb_pipeline_run_dev(p, d_x, x_length, d_t, d_f, d_sp, d_ap, d_y, y_length, st);
    p->host_out = wb_pipeline::HostOut();
    if (rc) return rc;
    WB_CUDA_CHECK(cudaStreamSynchronize(st));
    return p->ws.read_error_flag(st);
  }
  if ((rc = wb_pipeline_run_dev(p, d_x, x_length, d_t, d_f, d_sp, d_ap, d_y, y_length, st))) return rc;
  if (tpos) WB_CUDA_CHECK(cudaMemcpyAsync(tpos, d_t, sizeof(double) * f0_length, cudaMemcpyDeviceToHost, st));
  if (f0) WB_CUDA_CHECK(cudaMemcpyAsync(f0, d_f, sizeof(double) * f0_length, cudaMemcpyDeviceToHost, st));
  if (sp) WB_CUDA_CHECK(cudaMemcpyAsync(sp, d_sp, sizeof(double) * (size_t)f0_length * bins, cudaMemcpyDeviceToHost, st));
  if (ap) WB_CUDA_CHECK(cudaMemcpyAsync(ap, d_ap, sizeof(double) * (size_t)f0_length * bins, cudaMemcpyDeviceToHost, st));
  if (y_length > 0) WB_CUDA_CHECK(cudaMemcpyAsync(y, d_y, sizeof(double) * (size_t)y_length, cudaMemcpyDeviceToHost, st));
  WB_CUDA_CHECK(cudaStreamSynchronize(st));
  return p->ws.read_error_flag(st);
}

// ---- one long stream sharded over ranks (BASELINE configs[3]) ----------------------
// The frames of a stream are independent given the whole stream's f0, EXCEPT for their position in the
// process-global randn() stream (a prefix sum over all earlier frames and stages) and, in Synthesis, the
// sequential phase sum that places the pulses.  Both are cheap functions of f0 (and of the Love Train
// decisions), so every rank recomputes them for the whole stream and does the heavy per-frame / per-pulse work
// for its own range only; the results equal an unsharded run bit for bit.  Call order per rank:
//   begin -> envelope (CheapTrick + Love Train rows) -> [all-gather ap0] -> aperiodicity (D4C body rows)
//   -> synthesis (sample range).  The collectives themselves are the caller's (torch.distributed / NCCL).
namespace {
__global__ void frame_times_kernel(double *__restrict__ tpos, int n, double frame_period) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tpos[i] = i * frame_period / 1000.0;   // harvest.cpp:196-199
}
}  // namespace

int wb_pipeline_stream_begin_dev(wb_pipeline_t *p, const double *d_f0_all, int f0_length, int out_length, void *stream_) {
  return wb_pipeline_stream_begin_range_dev(p, d_f0_all, f0_length, out_length, 0, -1, stream_);
}

int wb_pipeline_stream_begin_range_dev(wb_pipeline_t *p, const double *d_f0_all, int f0_length, int out_length,
                                       int sample_begin, int sample_end, void *stream_) {
  if (!p || !d_f0_all || f0_length < 2 || out_length < 0) return WB_ERR_ARG;
  if (sample_end >= 0 && (sample_begin < 0 || sample_begin > sample_end || sample_end > out_length)) return WB_ERR_ARG;
  cudaStream_t stream = pick_stream(stream_);
  const double fp = p->plan.opt.frame_period;
  double *d_tpos = (double *)p->ws.get("st_tpos", sizeof(double) * f0_length);
  unsigned long long *rng_pos = (unsigned long long *)p->ws.get("pl_rng_pos", sizeof(unsigned long long) * 4);
  if (!d_tpos || !rng_pos) return WB_ERR_CUDA;
  WB_LAUNCH("frame_times_kernel", frame_times_kernel<<<(f0_length + 255) / 256, 256, 0, stream>>>(d_tpos, f0_length, fp));
  WB_CUDA_CHECK(cudaGetLastError());
  p->stream_f0_length = f0_length;
  if (p->private_rng)
    WB_CUDA_CHECK(cudaMemcpyAsync(p->d_rng_private, p->d_rng_seed, sizeof(WbRngState), cudaMemcpyDeviceToDevice, stream));
  // the time base / exact phase scan / pulse list of the whole stream depend on f0 only: side stream
  if (out_length > 0) {
    WB_CUDA_CHECK(cudaEventRecord(p->ev_f0, stream));
    WB_CUDA_CHECK(cudaStreamWaitEvent(p->side, p->ev_f0, 0));
    int rc = wb_synthesis_timebase(&p->ws, p->fs, p->ct.fft_size, fp, d_f0_all, f0_length, out_length, p->side, nullptr,
                                   sample_begin, sample_end);
    if (rc) return rc;
    WB_CUDA_CHECK(cudaEventRecord(p->ev_tb, p->side));
  }
  p->stream_samples[0] = sample_end >= 0 ? sample_begin : 0;
  p->stream_samples[1] = sample_end >= 0 ? sample_end : out_length;
  return WB_OK;
}

int wb_pipeline_stream_envelope_dev(wb_pipeline_t *p, const double *d_x, int x_length, const double *d_f0_all,
                                    int f0_length, int frame_begin, int frame_end, double *d_sp_rows,
                                    double *d_ap0_all, void *stream_) {
  if (!p || !d_x || !d_f0_all || !d_sp_rows || !d_ap0_all || x_length <= 0) return WB_ERR_ARG;
  cudaStream_t stream = pick_stream(stream_);
  double *d_tpos = (double *)p->ws.find("st_tpos");
  unsigned long long *rng_pos = (unsigned long long *)p->ws.find("pl_rng_pos");
  if (!d_tpos || !rng_pos || f0_length != p->stream_f0_length) return WB_ERR_ARG;   // begin has not been called (with this length)
  if (frame_begin < 0 || frame_end > f0_length || frame_begin > frame_end) return WB_ERR_ARG;
  WbRngState *rng = p->private_rng ? p->d_rng_private : wb_rng_global_state();
  const WbFrameRange range = {frame_begin, frame_end};
  WbRngCursor c_ct;
  c_ct.state = rng; c_ct.skip_out = rng_pos + 0; c_ct.advance = false;
  int rc;
  if ((rc = wb_cheaptrick_run(&p->ws, p->fs, p->ct.fft_size, p->ct.q1, p->ct_f0_floor_internal, d_x, x_length, d_tpos,
                              d_f0_all, f0_length, d_sp_rows, c_ct, stream, nullptr, &range)))
    return rc;
  WbRngCursor c_lt;
  c_lt.state = rng; c_lt.skip_in = rng_pos + 0; c_lt.advance = false;
  return wb_d4c_run(&p->ws, p->fs, p->d4c.threshold, d_x, x_length, d_tpos, d_f0_all, f0_length, p->ct.fft_size, nullptr,
                    c_lt, stream, nullptr, &range, 1, d_ap0_all);
}

// The two halves of the envelope call, for callers that want the Love Train decisions (8 bytes per frame, but
// the all-gather that distributes them waits for the slowest rank) on their way BEFORE the longer CheapTrick
// work starts: Love Train only needs CheapTrick's draw COUNT for its position in the randn() stream.
int wb_pipeline_stream_lovetrain_dev(wb_pipeline_t *p, const double *d_x, int x_length, const double *d_f0_all,
                                     int f0_length, int frame_begin, int frame_end, double *d_ap0_all, void *stream_) {
  if (!p || !d_x || !d_f0_all || !d_ap0_all || x_length <= 0) return WB_ERR_ARG;
  cudaStream_t stream = pick_stream(stream_);
  double *d_tpos = (double *)p->ws.find("st_tpos");
  unsigned long long *rng_pos = (unsigned long long *)p->ws.find("pl_rng_pos");
  if (!d_tpos || !rng_pos || f0_length != p->stream_f0_length) return WB_ERR_ARG;   // begin has not been called (with this length)
  if (frame_begin < 0 || frame_end > f0_length || frame_begin > frame_end) return WB_ERR_ARG;
  WbRngState *rng = p->private_rng ? p->d_rng_private : wb_rng_global_state();
  // CheapTrick's count of the whole stream -> rng_pos[0] (an empty row range: bookkeeping only)
  const WbFrameRange none = {frame_begin, frame_begin};
  WbRngCursor c_ct;
  c_ct.state = rng; c_ct.skip_out = rng_pos + 0; c_ct.advance = false;
  int rc;
  if ((rc = wb_cheaptrick_run(&p->ws, p->fs, p->ct.fft_size, p->ct.q1, p->ct_f0_floor_internal, d_x, x_length, d_tpos,
                              d_f0_all, f0_length, nullptr, c_ct, stream, nullptr, &none)))
    return rc;
  const WbFrameRange range = {frame_begin, frame_end};
  WbRngCursor c_lt;
  c_lt.state = rng; c_lt.skip_in = rng_pos + 0; c_lt.advance = false;
  return wb_d4c_run(&p->ws, p->fs, p->d4c.threshold, d_x, x_length, d_tpos, d_f0_all, f0_length, p->ct.fft_size, nullptr,
                    c_lt, stream, nullptr, &range, 1, d_ap0_all);
}

int wb_pipeline_stream_cheaptrick_dev(wb_pipeline_t *p, const double *d_x, int x_length, const double *d_f0_all,
                                      int f0_length, int frame_begin, int frame_end, double *d_sp_rows, void *stream_) {
  if (!p || !d_x || !d_f0_all || !d_sp_rows || x_length <= 0) return WB_ERR_ARG;
  cudaStream_t stream = pick_stream(stream_);
  double *d_tpos = (double *)p->ws.find("st_tpos");
  unsigned long long *rng_pos = (unsigned long long *)p->ws.find("pl_rng_pos");
  if (!d_tpos || !rng_pos || f0_length != p->stream_f0_length) return WB_ERR_ARG;
  if (frame_begin < 0 || frame_end > f0_length || frame_begin > frame_end) return WB_ERR_ARG;
  WbRngState *rng = p->private_rng ? p->d_rng_private : wb_rng_global_state();
  const WbFrameRange range = {frame_begin, frame_end};
  WbRngCursor c_ct;
  c_ct.state = rng; c_ct.skip_out = rng_pos + 0; c_ct.advance = false;
  return wb_cheaptrick_run(&p->ws, p->fs, p->ct.fft_size, p->ct.q1, p->ct_f0_floor_internal, d_x, x_length, d_tpos,
                           d_f0_all, f0_length, d_sp_rows, c_ct, stream, nullptr, &range);
}

int wb_pipeline_stream_aperiodicity_dev(wb_pipeline_t *p, const double *d_x, int x_length, const double *d_f0_all,
                                        const double *d_ap0_all, int f0_length, int frame_begin, int frame_end,
                                        double *d_ap_rows, void *stream_) {
  if (!p || !d_x || !d_f0_all || !d_ap_rows || !d_ap0_all || x_length <= 0) return WB_ERR_ARG;
  cudaStream_t stream = pick_stream(stream_);
  double *d_tpos = (double *)p->ws.find("st_tpos");
  unsigned long long *rng_pos = (unsigned long long *)p->ws.find("pl_rng_pos");
  if (!d_tpos || !rng_pos || f0_length != p->stream_f0_length) return WB_ERR_ARG;
  if (frame_begin < 0 || frame_end > f0_length || frame_begin > frame_end) return WB_ERR_ARG;
  WbRngState *rng = p->private_rng ? p->d_rng_private : wb_rng_global_state();
  const WbFrameRange range = {frame_begin, frame_end};
  WbRngCursor c;
  c.state = rng; c.skip_in = rng_pos + 0; c.skip_out = rng_pos + 1; c.advance = false;
  return wb_d4c_run(&p->ws, p->fs, p->d4c.threshold, d_x, x_length, d_tpos, d_f0_all, f0_length, p->ct.fft_size, d_ap_rows,
                    c, stream, nullptr, &range, 2, const_cast<double *>(d_ap0_all));
}

int wb_pipeline_stream_synthesis_dev(wb_pipeline_t *p, int f0_length, const double *d_sp_rows, const double *d_ap_rows,
                                     int row_begin, int n_rows, int out_length, int sample_begin, int sample_end,
                                     double *d_out, void *stream_) {
  if (!p || !d_sp_rows || !d_ap_rows || !d_out || out_length <= 0) return WB_ERR_ARG;
  cudaStream_t stream = pick_stream(stream_);
  unsigned long long *rng_pos = (unsigned long long *)p->ws.find("pl_rng_pos");
  if (!rng_pos || !p->ws.find("syn_pidx") || f0_length != p->stream_f0_length) return WB_ERR_ARG;   // begin (with out_length > 0) has not been called
  if (sample_begin < p->stream_samples[0] || sample_end > p->stream_samples[1]) return WB_ERR_ARG;     // outside the range begin was given
  WbRngState *rng = p->private_rng ? p->d_rng_private : wb_rng_global_state();
  WB_CUDA_CHECK(cudaStreamWaitEvent(stream, p->ev_tb, 0));
  WbRngCursor c;
  c.state = rng; c.skip_in = rng_pos + 1; c.advance = false;  // (several ranges may follow: wb_pipeline_stream_end_dev moves the state)
  return wb_synthesis_render_range(&p->ws, p->fs, p->ct.fft_size, p->plan.opt.frame_period, f0_length, d_sp_rows, d_ap_rows,
                                   row_begin, n_rows, out_length, sample_begin, sample_end, d_out,
                                   p->stream_f0_bound > 0.0 ? p->stream_f0_bound : p->plan.opt.f0_ceil * 1.25, c, stream);
}

// moves the randn() state past the whole stream's draws (what one reference process would leave behind)
int wb_pipeline_stream_end_dev(wb_pipeline_t *p, void *stream_) {
  if (!p) return WB_ERR_ARG;
  cudaStream_t stream = pick_stream(stream_);
  unsigned long long *rng_pos = (unsigned long long *)p->ws.find("pl_rng_pos");
  unsigned long long *d_ncount = (unsigned long long *)p->ws.find("syn_ncount");
  if (!rng_pos) return WB_ERR_ARG;
  WbRngState *rng = p->private_rng ? p->d_rng_private : wb_rng_global_state();
  return wb_rng_advance(rng, rng_pos + 1, d_ncount, stream);
}

// ---- streaming synthesis (the real-time use the demo points at, test/test.cpp:353-358) -----------------------
struct wb_synthesis_stream { WbSynStream *impl; };

int wb_synthesis_stream_create(int fs, int fft_size, double frame_period_ms, double f0_upper_bound,
                               wb_synthesis_stream_t **out) {
  if (!out) return WB_ERR_ARG;
  int rc = ctx_init();
  if (rc) return rc;
  WbSynStream *impl = wb_synstream_create(fs, fft_size, frame_period_ms, f0_upper_bound);
  if (!impl) return WB_ERR_UNSUPPORTED;
  wb_synthesis_stream *h = new (std::nothrow) wb_synthesis_stream();
  if (!h) { wb_synstream_destroy(impl); return WB_ERR_ARG; }
  h->impl = impl;
  *out = h;
  return WB_OK;
}
void wb_synthesis_stream_destroy(wb_synthesis_stream_t *s) {
  if (!s) return;
  wb_synstream_destroy(s->impl);
  delete s;
}
int wb_synthesis_stream_push(wb_synthesis_stream_t *s, const double *f0, const double *spectrogram, const double *aperiodicity,
                             int n_frames, double *out, int out_capacity, int *n_out) {
  if (!s) return WB_ERR_ARG;
  return wb_synstream_push(s->impl, f0, spectrogram, aperiodicity, n_frames, out, out_capacity, n_out, g_stream);
}
int wb_synthesis_stream_finish(wb_synthesis_stream_t *s, int out_length_total, double *out, int out_capacity, int *n_out) {
  if (!s) return WB_ERR_ARG;
  return wb_synstream_finish(s->impl, out_length_total, out, out_capacity, n_out, g_stream);
}

// wav in -> wav out (test/test.cpp:288-384 with tools/audioio.cpp either side): 16-bit PCM crosses PCIe, the
// sample-format conversions of wavread / wavwrite run on the device (wb_io.cu)
int wb_pipeline_run_pcm16(wb_pipeline_t *p, const short *pcm_in, int x_length, short *pcm_out, int y_length) {
  if (!p || !pcm_in || x_length <= 0 || y_length < 0 || (y_length > 0 && !pcm_out)) return WB_ERR_ARG;
  cudaStream_t st = g_stream;
  short *d_pcm_in = (short *)p->ws.get("pl_pcm_in", sizeof(short) * (size_t)x_length);
  short *d_pcm_out = (short *)p->ws.get("pl_pcm_out", sizeof(short) * (size_t)(y_length > 0 ? y_length : 1));
  double *d_x = (double *)p->ws.get("h_x", sizeof(double) * (size_t)x_length);
  double *d_y = (double *)p->ws.get("pl_y", sizeof(double) * (size_t)(y_length > 0 ? y_length : 1));
  if (!d_pcm_in || !d_pcm_out || !d_x || !d_y) return WB_ERR_CUDA;
  int rc;
  WB_CUDA_CHECK(cudaMemcpyAsync(d_pcm_in, pcm_in, sizeof(short) * (size_t)x_length, cudaMemcpyHostToDevice, st));
  if ((rc = wb_pcm16_to_f64_run(d_pcm_in, x_length, d_x, st))) return rc;
  if ((rc = wb_pipeline_run_dev(p, d_x, x_length, nullptr, nullptr, nullptr, nullptr, d_y, y_length, st))) return rc;
  if (y_length > 0) {
    if ((rc = wb_f64_to_pcm16_run(d_y, y_length, d_pcm_out, st))) return rc;
    WB_CUDA_CHECK(cudaMemcpyAsync(pcm_out, d_pcm_out, sizeof(short) * (size_t)y_length, cudaMemcpyDeviceToHost, st));
  }
  WB_CUDA_CHECK(cudaStreamSynchronize(st));
  return p->ws.read_error_flag(st);
}

// analysis (+ optional re-synthesis) with the outputs narrowed to fp32 on the device: half the bytes back to
// the host, contiguous [f0_length][fft_size/2+1] matrices.  Arithmetic stays fp64.
int wb_pipeline_run_f32(wb_pipeline_t *p, const double *x, int x_length, float *f0, float *sp, float *ap, float *y,
                        int y_length) {
  if (!p || !x || x_length <= 0 || y_length < 0 || (y_length > 0 && !y)) return WB_ERR_ARG;
  cudaStream_t st = g_stream;
  const int f0_length = wb_pipeline_f0_length(p, x_length);
  const size_t bins = p->ct.fft_size / 2 + 1, cells = (size_t)f0_length * bins;
  double *d_x;
  int rc;
  if ((rc = vec_to_device(&p->ws, "h_x", x, x_length, &d_x, st))) return rc;
  double *d_f = (double *)p->ws.get("pl_f0", sizeof(double) * f0_length);
  double *d_sp = (double *)p->ws.get("pl_sp", sizeof(double) * cells);
  double *d_ap = (double *)p->ws.get("pl_ap", sizeof(double) * cells);
  double *d_y = (double *)p->ws.get("pl_y", sizeof(double) * (size_t)(y_length > 0 ? y_length : 1));
  float *d_n = (float *)p->ws.get("pl_f32", sizeof(float) * (2 * cells + (size_t)f0_length + (size_t)y_length + 4));
  if (!d_f || !d_sp || !d_ap || !d_y || !d_n) return WB_ERR_CUDA;
  if ((rc = wb_pipeline_run_dev(p, d_x, x_length, nullptr, d_f, d_sp, d_ap, d_y, y_length, st))) return rc;
  float *n_sp = d_n, *n_ap = d_n + cells, *n_f0 = d_n + 2 * cells, *n_y = n_f0 + f0_length;
  if (sp) {
    if ((rc = wb_f64_to_f32_run(d_sp, cells, n_sp, st))) return rc;
    WB_CUDA_CHECK(cudaMemcpyAsync(sp, n_sp, sizeof(float) * cells, cudaMemcpyDeviceToHost, st));
  }
  if (ap) {
    if ((rc = wb_f64_to_f32_run(d_ap, cells, n_ap, st))) return rc;
    WB_CUDA_CHECK(cudaMemcpyAsync(ap, n_ap, sizeof(float) * cells, cudaMemcpyDeviceToHost, st));
  }
  if (f0) {
    if ((rc = wb_f64_to_f32_run(d_f, f0_length, n_f0, st))) return rc;
    WB_CUDA_CHECK(cudaMemcpyAsync(f0, n_f0, sizeof(float) * f0_length, cudaMemcpyDeviceToHost, st));
  }
  if (y_length > 0) {
    if ((rc = wb_f64_to_f32_run(d_y, y_length, n_y, st))) return rc;
    WB_CUDA_CHECK(cudaMemcpyAsync(y, n_y, sizeof(float) * (size_t)y_length, cudaMemcpyDeviceToHost, st));
  }
  WB_CUDA_CHECK(cudaStreamSynchronize(st));
  return p->ws.read_error_flag(st);
}

/* device-pointer conversions (asynchronous on `stream`) */
int wb_pcm16_to_f64_dev(const short *d_pcm, int n, double *d_x, void *stream) {
  int rc = ctx_init();
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!d_pcm || !d_x))) return WB_ERR_ARG;
  return wb_pcm16_to_f64_run(d_pcm, n, d_x, pick_stream(stream));
}
int wb_f64_to_pcm16_dev(const double *d_x, int n, short *d_pcm, void *stream) {
  int rc = ctx_init();
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!d_pcm || !d_x))) return WB_ERR_ARG;
  return wb_f64_to_pcm16_run(d_x, n, d_pcm, pick_stream(stream));
}
int wb_f64_to_f32_dev(const double *d_in, unsigned long long n, float *d_out, void *stream) {
  int rc = ctx_init();
  if (rc) return rc;
  if (n > 0 && (!d_in || !d_out)) return WB_ERR_ARG;
  return wb_f64_to_f32_run(d_in, (size_t)n, d_out, pick_stream(stream));
}

// ---- parameter modification (test/test.cpp:201-243) ------------------------------------------
int wb_parameter_modification_dev(double *d_f0, int f0_length, double *d_spectrogram, int fs, int fft_size,
                                  double f0_shift, double ratio, void *stream) {
  int rc = ctx_init();
  if (rc) return rc;
  if (f0_length < 0 || fs <= 0 || fft_size < 4) return WB_ERR_ARG;
  return wb_parameter_modification_run(d_f0, d_f0, f0_length, d_spectrogram, fs, fft_size, f0_shift, ratio,
                                       pick_stream(stream));
}

int wb_parameter_modification(double *f0, int f0_length, double **spectrogram, int fs, int fft_size,
                              double f0_shift, double ratio) {
  int rc = ctx_init();
  if (rc) return rc;
  if (f0_length < 0 || fs <= 0 || fft_size < 4) return WB_ERR_ARG;
  if (f0_length == 0) return WB_OK;
  static WbWorkspace ws;   // a free function in the reference's demo: one shared workspace ...
  static std::mutex ws_mutex;   // ... so concurrent callers take turns (the reference's function is stateless)
  std::lock_guard<std::mutex> lock(ws_mutex);
  cudaStream_t st = g_stream;
  const int bins = fft_size / 2 + 1;
  const bool do_f0 = f0 && f0_shift == f0_shift, do_sp = spectrogram && ratio > 0.0;
  double *d_f0 = nullptr, *d_sp = nullptr;
  if (do_f0 && (rc = vec_to_device(&ws, "mod_f0", f0, f0_length, &d_f0, st))) return rc;
  if (do_sp) {
    d_sp = (double *)ws.get("mod_sp", sizeof(double) * (size_t)f0_length * bins);
    if (!d_sp) return WB_ERR_CUDA;
    if ((rc = rows_to_device(&ws, "rows_stage_in", spectrogram, f0_length, bins, d_sp, st))) return rc;
  }
  if ((rc = wb_parameter_modification_run(d_f0, d_f0, f0_length, d_sp, fs, fft_size, do_f0 ? f0_shift : NAN,
                                          do_sp ? ratio : 0.0, st)))
    return rc;
  if (do_f0) WB_CUDA_CHECK(cudaMemcpyAsync(f0, d_f0, sizeof(double) * f0_length, cudaMemcpyDeviceToHost, st));
  if (do_sp && (rc = rows_to_host(&ws, d_sp, f0_length, bins, spectrogram, st))) return rc;
  WB_CUDA_CHECK(cudaStreamSynchronize(st));
  return WB_OK;
}

// ---- codec (include/codec.hpp:23-88) --------------------------------------------------------
static WbWorkspace *codec_ws() {
  static WbWorkspace ws;  // codec functions are free functions in the reference: one shared workspace
  return &ws;
}

static int codec_rows(int kind, const double *const *in, int in_cols, int f0_length, int fs, int fft_size, int nd,
                      double **out, int out_cols) {
  int rc = ctx_init();
  if (rc) return rc;
  if (!in || !out || f0_length < 0 || fs <= 0 || fft_size < 2) return WB_ERR_ARG;
  if (f0_length == 0) return WB_OK;
  if (in_cols <= 0 || out_cols <= 0) return WB_ERR_UNSUPPORTED;
  WbWorkspace *ws = codec_ws();
  static std::mutex ws_mutex;   // the reference's codec functions are stateless: concurrent callers take turns on the staging buffers
  std::lock_guard<std::mutex> lock(ws_mutex);
  cudaStream_t st = g_stream;
  double *d_in = (double *)ws->get("codec_in", sizeof(double) * (size_t)f0_length * in_cols);
  double *d_out = (double *)ws->get("codec_out", sizeof(double) * (size_t)f0_length * out_cols);
  if (!d_in || !d_out) return WB_ERR_CUDA;
  if ((rc = rows_to_device(ws, "rows_stage_in", in, f0_length, in_cols, d_in, st))) return rc;
  switch (kind) {
    case 0: rc = wb_code_aperiodicity_dev(d_in, f0_length, fs, fft_size, d_out, st); break;
    case 1: rc = wb_decode_aperiodicity_dev(d_in, f0_length, fs, fft_size, d_out, st); break;
    case 2: rc = wb_code_spectral_envelope_dev(ws, d_in, f0_length, fs, fft_size, nd, d_out, st); break;
    default: rc = wb_decode_spectral_envelope_dev(ws, d_in, f0_length, fs, fft_size, nd, d_out, st); break;
  }
  if (rc) return rc;
  return rows_to_host(ws, d_out, f0_length, out_cols, out, st);
}

int wb_code_aperiodicity(const double *const *aperiodicity, int f0_length, int fs, int fft_size,
                         double **coded_aperiodicity) {
  return codec_rows(0, aperiodicity, fft_size / 2 + 1, f0_length, fs, fft_size, 0, coded_aperiodicity,
                    wb_number_of_aperiodicities(fs));
}
int wb_decode_aperiodicity(const double *const *coded_aperiodicity, int f0_length, int fs, int fft_size,
                           double **aperiodicity) {
  return codec_rows(1, coded_aperiodicity, wb_number_of_aperiodicities(fs), f0_length, fs, fft_size, 0, aperiodicity,
                    fft_size / 2 + 1);
}
int wb_code_spectral_envelope(const double *const *spectrogram, int f0_length, int fs, int fft_size,
                              int number_of_dimensions, double **coded_spectral_envelope) {
  return codec_rows(2, spectrogram, fft_size / 2 + 1, f0_length, fs, fft_size, number_of_dimensions,
                    coded_spectral_envelope, number_of_dimensions);
}
int wb_decode_spectral_envelope(const double *const *coded_spectral_envelope, int f0_length, int fs, int fft_size,
                                int number_of_dimensions, double **spectrogram) {
  return codec_rows(3, coded_spectral_envelope, number_of_dimensions, f0_length, fs, fft_size, number_of_dimensions,
                    spectrogram, fft_size / 2 + 1);
}

// contiguous device-pointer variants; asynchronous on `stream`
int wb_codec_dev(int kind, const double *d_in, int f0_length, int fs, int fft_size, int number_of_dimensions,
                 double *d_out, void *stream) {
  int rc = ctx_init();
  if (rc) return rc;
  if (!d_in || !d_out || f0_length < 0) return WB_ERR_ARG;
  cudaStream_t st = pick_stream(stream);
  switch (kind) {
    case 0: return wb_code_aperiodicity_dev(d_in, f0_length, fs, fft_size, d_out, st);
    case 1: return wb_decode_aperiodicity_dev(d_in, f0_length, fs, fft_size, d_out, st);
    case 2: return wb_code_spectral_envelope_dev(codec_ws(), d_in, f0_length, fs, fft_size, number_of_dimensions, d_out, st);
    case 3: return wb_decode_spectral_envelope_dev(codec_ws(), d_in, f0_length, fs, fft_size, number_of_dimensions, d_out, st);
    default: return WB_ERR_ARG;
  }
}

/* test / bench hook: copies n_bytes of a named internal device buffer of the last run */
int wb_pipeline_debug_read(wb_pipeline_t *p, const char *name, void *out, unsigned long long n_bytes) {
  if (!p || !name || !out) return WB_ERR_ARG;
  size_t have = 0;
  void *d = p->ws.find(name, &have);
  if (!d || n_bytes > have) return WB_ERR_ARG;   // unknown buffer, or more than it holds
  WB_CUDA_CHECK(cudaDeviceSynchronize());
  WB_CUDA_CHECK(cudaMemcpy(out, d, n_bytes, cudaMemcpyDeviceToHost));
  return WB_OK;
}

// ---- errors detected on the device ---------------------------------------------------------------
// Kernels flag conditions they cannot handle (more pulses than the f0 bound allows for, a smoothing width beyond
// the scratch capacity) in a per-handle device word.  The host-pointer entry points read it before they return;
// the asynchronous device-pointer entry points cannot, so their callers ask here once they have synchronised
// (or want to): waits for `stream`, returns WB_OK or the flagged status, and clears the flag.
int wb_harvest_last_error(wb_harvest_t *h, void *stream) { return h ? h->ws.read_error_flag(pick_stream(stream)) : WB_ERR_ARG; }
int wb_cheaptrick_last_error(wb_cheaptrick_t *h, void *stream) { return h ? h->ws.read_error_flag(pick_stream(stream)) : WB_ERR_ARG; }
int wb_d4c_last_error(wb_d4c_t *h, void *stream) { return h ? h->ws.read_error_flag(pick_stream(stream)) : WB_ERR_ARG; }
int wb_synthesis_last_error(wb_synthesis_t *h, void *stream) { return h ? h->ws.read_error_flag(pick_stream(stream)) : WB_ERR_ARG; }
int wb_pipeline_last_error(wb_pipeline_t *p, void *stream) { return p ? p->ws.read_error_flag(pick_stream(stream)) : WB_ERR_ARG; }

/* Sharded streams whose f0 contour does not come from Harvest (wb_pipeline_stream_begin_dev with an external
 * contour): upper bound of its values, which sizes the pulse buffers of the synthesis calls.  <= 0: Harvest's
 * f0_ceil * 1.25 (the default). */
int wb_pipeline_set_stream_f0_bound(wb_pipeline_t *p, double f0_upper_bound) {
  if (!p || f0_upper_bound != f0_upper_bound) return WB_ERR_ARG;
  p->stream_f0_bound = f0_upper_bound;
  return WB_OK;
}

// ---- measurement hooks ---------------------------------------------------------------------
int wb_measure_fp64_peak(double *tflops) {
  int rc = ctx_init();
  if (rc) return rc;
  return wb_measure_fp64_peak_tflops(tflops);
}

unsigned long long wb_launch_count(void) { return wb_launch_counter(); }
void *wb_stream(void) { return ctx_init() ? nullptr : (void *)g_stream; }
void wb_profile_enable(int on) { wb_prof_set_enabled(on); }
void wb_profile_reset(void) { wb_prof_reset(); }
int wb_profile_collect(void) { return wb_prof_collect(); }
int wb_profile_query(const char *kernel_name, double *total_ms, int *count) { return wb_prof_query(kernel_name, total_ms, count); }
int wb_profile_names(char *buf, int buf_len) { return wb_prof_names(buf, buf_len); }

}  // extern "C"
