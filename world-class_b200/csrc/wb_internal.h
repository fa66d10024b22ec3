// Internal host-side plumbing shared by the stage files (not part of the C-ABI).
#pragma once
#include <map>
#include <string>

#include "wb_common.cuh"
#include "wb_rng.cuh"

// Grow-only named device scratch buffers (one workspace per stage handle / pipeline).
class WbWorkspace {
 public:
  WbWorkspace();
  ~WbWorkspace();
  void *get(const std::string &name, size_t bytes);        // device memory (allocates / grows)
  // look-up only: the buffer and its size if `name` exists, else null (never allocates)
  void *find(const std::string &name, size_t *bytes_out = nullptr) const;
  // bumped whenever a device buffer is freed (a buffer had to grow): captured CUDA graphs that hold raw
  // pointers into the workspace are stale once this differs from the value at capture time
  unsigned long long generation() const { return generation_; }
  // like get(), but the first `keep_bytes` of an existing buffer survive a reallocation (synchronises `stream`)
  void *get_keep(const std::string &name, size_t bytes, size_t keep_bytes, cudaStream_t stream);
  void *get_pinned(const std::string &name, size_t bytes);  // page-locked host memory
  int *error_flag();                                        // mapped page-locked int (device-writable, host-readable), starts at 0
  int read_error_flag(cudaStream_t stream);                 // syncs the stream
 private:
  struct Buf { void *p; size_t bytes; };
  std::map<std::string, Buf> dev_, pinned_;
  unsigned long long generation_;
  int *d_err_;
};

// number of SMs of the current device (queried once; grids of persistent kernels are multiples of it)
int wb_sm_count();

// e^{+2 pi i k / n}, k = 0..n-1, on the device; cached per n (power of two).
const cplx *wb_twiddle_table(int n);

// offsets[0..n] <- exclusive prefix sums of counts[0..n) (offsets[n] = total); optionally publishes
// *d_skip_out = *d_skip_in + total (randn stream bookkeeping, see WbRngCursor)
// (ws: scratch for the grid-wide version used above 16384 items; null = always one CTA)
int wb_exclusive_scan_u64(const unsigned long long *d_counts, unsigned long long *d_offsets, int n,
                          cudaStream_t stream, const unsigned long long *d_skip_in = nullptr,
                          unsigned long long *d_skip_out = nullptr, WbWorkspace *ws = nullptr);

// Host-pointer entry points overlap the device -> host transfer of finished output rows with the frames
// still being computed: the frame kernel is launched in `n` row ranges [bounds[c], bounds[c+1]) that
// alternate between the caller's stream and `alt` (they are independent, so they overlap each other),
// and ev[c] is recorded when range c is complete.  Null = one launch.
struct WbRowChunks {
  int n = 0;
  int bounds[17] = {0};
  cudaEvent_t ev[16] = {nullptr};
  cudaStream_t alt = nullptr;
  cudaEvent_t ev_ready = nullptr;   // recorded on the caller's stream when the frame kernel's inputs are ready
};

// A contiguous range of the frames of one long stream (frames shard across ranks).
// The per-frame stages take the WHOLE stream's f0 (the randn() position of a frame is a prefix sum over all
// earlier frames) but compute, draw noise for and write only rows [begin, end); the output pointer then
// addresses row `begin`.  The cursor bookkeeping (skip_out, advance) is that of the whole stream.
struct WbFrameRange { int begin; int end; };
// rel[i] = offsets[i] - offsets[begin] for i in [begin, end]; *skip_range = *skip_in + offsets[begin];
// *count_range = offsets[end] - offsets[begin]
int wb_range_offsets(const unsigned long long *d_offsets, WbFrameRange range, unsigned long long *d_rel,
                     const unsigned long long *d_skip_in, unsigned long long *d_skip_range,
                     unsigned long long *d_count_range, cudaStream_t stream);

int wb_cheaptrick_run(WbWorkspace *ws, int fs, int fft_size, double q1, double f0_floor_internal,
                      const double *d_x, int x_length, const double *d_tpos, const double *d_f0,
                      int f0_length, double *d_sp, const WbRngCursor &rng, cudaStream_t stream,
                      const WbRowChunks *chunks = nullptr, const WbFrameRange *range = nullptr);

// stand-alone batched transforms (wb_fftapi.cu); kind 0 r2c, 1 c2r, 2 c2c fwd, 3 c2c bwd
int wb_fft_batch_dev(int kind, const void *d_in, int n, int batch, void *d_out, cudaStream_t stream);

// Two halves of a stage's rows on two streams: the randn() fill of the second half runs beside the frame kernel of
// the first (integer work beside fp64 work) instead of in front of it.  `alt` forks from and joins the caller's stream.
struct WbStageSplit {
  cudaStream_t alt = nullptr;
  cudaEvent_t fork[2] = {nullptr, nullptr}, join[2] = {nullptr, nullptr};   // [0] Love Train, [1] body
};

// D4C (wb_d4c.cu)
int wb_d4c_fft_size(int fs);
int wb_d4c_lt_fft_size(int fs);
int wb_number_of_aperiodicities(int fs);
// phase: 0 = Love Train + body (d_ap0_ext may be null: internal buffer); sharded streams run 1 = Love Train only
// (writes d_ap0_ext[begin, end)) and, once every rank's decisions are gathered, 2 = body only (reads all of
// d_ap0_ext).  With a range, d_ap addresses row `begin`.
int wb_d4c_run(WbWorkspace *ws, int fs, double threshold, const double *d_x, int x_length, const double *d_tpos,
               const double *d_f0, int f0_length, int out_fft_size, double *d_ap, const WbRngCursor &rng,
               cudaStream_t stream, const WbRowChunks *chunks = nullptr, const WbFrameRange *range = nullptr,
               int phase = 0, double *d_ap0_ext = nullptr, const WbStageSplit *split = nullptr);

// Synthesis (wb_synthesis.cu)
// Samples [sample_begin, sample_end) of the whole stream's waveform.  Must follow wb_synthesis_timebase (whole
// stream, without noise) on `ws`.  d_sp / d_ap hold rows [row_begin, row_begin + n_rows) and must cover every
// frame a pulse touching the sample range interpolates between; d_out addresses sample `sample_begin`.
int wb_synthesis_render_range(WbWorkspace *ws, int fs, int fft_size, double frame_period_ms, int f0_length,
                              const double *d_sp, const double *d_ap, int row_begin, int n_rows, int out_length,
                              int sample_begin, int sample_end, double *d_out, double f0_upper_bound,
                              const WbRngCursor &rng, cudaStream_t stream, int slot = 0, cudaEvent_t rows_ready = nullptr);
int wb_synthesis_prepare(WbWorkspace *ws, int fft_size, cudaStream_t stream);
int wb_synthesis_run(WbWorkspace *ws, int fs, int fft_size, double frame_period_ms, const double *d_f0,
                     int f0_length, const double *d_sp, const double *d_ap, int out_length, double *d_out,
                     double f0_upper_bound, const WbRngCursor &rng, cudaStream_t stream);
int wb_synthesis_timebase(WbWorkspace *ws, int fs, int fft_size, double frame_period_ms, const double *d_f0,
                          int f0_length, int out_length, cudaStream_t stream,
                          const WbRngCursor *noise_cursor = nullptr, int sample_begin = 0, int sample_end = -1);
int wb_synthesis_render(WbWorkspace *ws, int fs, int fft_size, double frame_period_ms, int f0_length,
                        const double *d_sp, const double *d_ap, int out_length, double *d_out,
                        double f0_upper_bound, const WbRngCursor &rng, cudaStream_t stream,
                        bool noise_ready = false);

// codec (wb_codec.cu)
int wb_code_aperiodicity_dev(const double *d_ap, int f0_length, int fs, int fft_size, double *d_coded,
                             cudaStream_t stream);
int wb_decode_aperiodicity_dev(const double *d_coded, int f0_length, int fs, int fft_size, double *d_ap,
                               cudaStream_t stream);
int wb_code_spectral_envelope_dev(WbWorkspace *ws, const double *d_sp, int f0_length, int fs, int fft_size, int nd,
                                  double *d_coded, cudaStream_t stream);
int wb_decode_spectral_envelope_dev(WbWorkspace *ws, const double *d_coded, int f0_length, int fs, int fft_size,
                                    int nd, double *d_sp, cudaStream_t stream);

// parameter modification (wb_modify.cu): test/test.cpp:201-243
int wb_parameter_modification_run(const double *d_f0_in, double *d_f0_out, int f0_length, double *d_sp, int fs,
                                  int fft_size, double f0_shift, double ratio, cudaStream_t stream);

// sample-format conversions (wb_io.cu): tools/audioio.cpp:176-180 (wavwrite), :232-249 (wavread, 16 bit); fp32 narrowing
int wb_pcm16_to_f64_run(const short *d_in, int n, double *d_out, cudaStream_t stream);
int wb_f64_to_pcm16_run(const double *d_in, int n, short *d_out, cudaStream_t stream);
int wb_f64_to_f32_run(const double *d_in, size_t n, float *d_out, cudaStream_t stream);

// Streaming synthesis: frames arrive in pieces, samples are emitted as soon as no
// later pulse can reach them; the concatenated output equals Synthesis::compute on all frames bit for bit
// (phase sum, pulse list and randn() positions are carried between pieces).  wb_synthesis.cu.
struct WbSynStream;
WbSynStream *wb_synstream_create(int fs, int fft_size, double frame_period_ms, double f0_upper_bound);
void wb_synstream_destroy(WbSynStream *s);
// HOST pointers; sp / ap contiguous [n_frames][fft_size/2+1]; emits at most out_capacity samples
int wb_synstream_push(WbSynStream *s, const double *f0, const double *sp, const double *ap, int n_frames, double *out,
                      int out_capacity, int *n_out, cudaStream_t stream);
// no more frames: emits the rest up to out_length_total (repeat until *n_out == 0 if out_capacity is short)
int wb_synstream_finish(WbSynStream *s, int out_length_total, double *out, int out_capacity, int *n_out, cudaStream_t stream);

// measured fp64 multiply-add throughput of the current device in TFLOP/s (wb_runtime.cu)
int wb_measure_fp64_peak_tflops(double *tflops_out);
