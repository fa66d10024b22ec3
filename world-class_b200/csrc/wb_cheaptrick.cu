// CheapTrick spectral envelope: one thread block per analysis frame, everything between
// the waveform gather and the final exp() stays in shared memory (3 FFTs per frame).
//
// Reference: /root/reference/src/cheaptrick.cpp
//   compute :48-95, generalBody :108-135, getWindowedWaveform :137-167,
//   setParametersForGetWindowedWaveform :169-196, getPowerSpectrum :198-218,
//   addInfinitesimalNoise :220-228, smoothingWithRecovery :230-276.
#include "wb_internal.h"
#include "wb_fft.cuh"
#include "wb_scan.cuh"
#include "wb_smooth.cuh"

namespace {

__device__ __forceinline__ double ct_current_f0(double f0, double f0_floor) {
  return (f0 <= f0_floor) ? WB_DEFAULT_F0 : f0;  // cheaptrick.cpp:76
}

// randn() calls made by frame i: window samples + one per bin (cheaptrick.cpp:153, :227); their exclusive
// prefix sums are the frames' positions in the stream
__global__ void __launch_bounds__(1024) ct_count_scan_kernel(const double *__restrict__ f0, int f0_length, int fs,
                                                             int fft_size, double f0_floor,
                                                             unsigned long long *__restrict__ offsets,
                                                             const unsigned long long *__restrict__ skip_in,
                                                             unsigned long long *__restrict__ skip_out) {
  wb_block_count_scan([&](int i) {
    const double cf0 = ct_current_f0(f0[i], f0_floor);
    const int hw = wb_round(1.5 * fs / cf0);
    return (unsigned long long)(2 * hw + 1) + (unsigned long long)(fft_size / 2 + 1);
  }, f0_length, offsets, skip_in, skip_out);
}

// the same count as a functor, for the grid-wide scan of long streams (wb_scan.cuh)
struct CtCountFn {
  const double *f0; int fs, fft_size; double f0_floor;
  __device__ unsigned long long operator()(int i) const {
    const double cf0 = ct_current_f0(f0[i], f0_floor);
    const int hw = wb_round(1.5 * fs / cf0);
    return (unsigned long long)(2 * hw + 1) + (unsigned long long)(fft_size / 2 + 1);
  }
};

struct CtParams {
  const double *x;
  int x_length;
  const double *tpos;
  const double *f0;
  int f0_length;
  int fs;
  int fft_size;
  int log2nc;
  double q1;
  double f0_floor;
  const cplx *twiddle;              // fft_size entries
  const double *noise;              // randn stream segment of this call
  const unsigned long long *noise_off;  // exclusive offsets per frame
  int noise_origin;                     // >= 0: the noise buffer starts at the draws of this frame (offsets relative to it)
  double *sp;                       // [f0_length][fft_size/2+1]
  int seg_capacity;
  int *error_flag;
  int frame_begin;                  // this launch covers frames frame_begin + blockIdx.x
};

template <int LOG2N>
__global__ void __launch_bounds__(256, 4) ct_frame_kernel(CtParams p) {
  extern __shared__ double2 smem_raw[];
  constexpr int N = 1 << LOG2N, NC = N / 2, bins = NC + 1;
  // one radix-8 butterfly per thread and pass (the launch uses min(256, N / 8) threads, at least 64): warp-local late passes
  constexpr bool WL = (NC / 8 <= 256);
  cplx *S = smem_raw;                                         // FFT slots
  double *A = reinterpret_cast<double *>(S + wb_fft_slots(NC));  // N + 2 doubles
  double *B = A + (N + 2);                                    // seg_capacity doubles
  double *red = B + p.seg_capacity;                           // 256 + 64 (one entry per thread for the scan, 64 for the block sums)
  double *W = reinterpret_cast<double *>(S);                  // packed real waveform view

  const int frame = p.frame_begin + blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const double f0 = ct_current_f0(p.f0[frame], p.f0_floor);
  const int fs = p.fs;
  const int hw = wb_round(1.5 * fs / f0);
  const int wlen = 2 * hw + 1;
  const double *noise = p.noise + (p.noise_off[frame] - (p.noise_origin >= 0 ? p.noise_off[p.noise_origin] : 0ull));

  // ---- F0-adaptive windowing (cheaptrick.cpp:137-196)
  const int origin = wb_round(p.tpos[frame] * fs + 0.001);
  double acc = 0.0;
  for (int j = tid; j < wlen; j += nt) {
    const double position = (j - hw) / 1.5 / fs;
    const double w = 0.5 * cos(WB_PI * position * f0) + 0.5;
    A[j] = w;
    acc += w * w;
  }
  const double average = sqrt(wb_block_sum(acc, red));
  double s1 = 0.0, s2 = 0.0;
  for (int j = tid; j < wlen; j += nt) {
    const double w = A[j] / average;
    A[j] = w;
    const int idx = wb_min_i(p.x_length - 1, wb_max_i(0, origin + j - hw));
    const double v = p.x[idx] * w + noise[j] * 0.000000000000001;
    W[wb_didx(j)] = v;
    s1 += v;
    s2 += w;
  }
  wb_block_sum2(s1, s2, red);
  const double coef = s1 / s2;
  for (int j = tid; j < N; j += nt) {
    if (j < wlen) W[wb_didx(j)] -= A[j] * coef;
    else W[wb_didx(j)] = 0.0;
  }
  __syncthreads();

  // ---- power spectrum (cheaptrick.cpp:198-218)
  wb_rfft_t<1, LOG2N - 1, WB_FFT_DEFAULT_RADIX, WL>(S, p.twiddle, [&](int k, cplx X) { A[k] = X.x * X.x + X.y * X.y; });
  wb_dc_correction(A, f0, fs, N);

  // ---- linear smoothing with width 2 f0 / 3 (cheaptrick.cpp:124-125)
  if (!wb_linear_smoothing(A, A, f0 * 2.0 / 3.0, fs, N, B, p.seg_capacity, red)) {
    if (tid == 0) atomicExch(p.error_flag, WB_ERR_UNSUPPORTED);
    return;
  }

  // ---- infinitesimal noise, log, mirror (cheaptrick.cpp:220-228, :255-258)
  for (int i = tid; i < bins; i += nt) {
    const double v = log(A[i] + fabs(noise[wlen + i]) * WB_EPS);
    W[wb_didx(i)] = v;
    if (i > 0 && i < NC) W[wb_didx(N - i)] = v;
  }
  __syncthreads();

  // ---- liftering in the cepstral domain (cheaptrick.cpp:238-269)
  const double q1 = p.q1;
  wb_rfft_t<1, LOG2N - 1, WB_FFT_DEFAULT_RADIX, WL>(S, p.twiddle, [&](int k, cplx X) {
    double sl = 1.0, cl = (1.0 - 2.0 * q1) + 2.0 * q1;
    if (k > 0) {
      const double quefrency = static_cast<double>(k) / fs;
      sl = sin(WB_PI * f0 * quefrency) / (WB_PI * f0 * quefrency);
      cl = (1.0 - 2.0 * q1) + 2.0 * q1 * cos(2.0 * WB_PI * quefrency * f0);
    }
    A[k] = X.x * sl * cl / N;
  });
  wb_irfft_t<-1, LOG2N - 1, WB_FFT_DEFAULT_RADIX, WL>(S, p.twiddle, [&](int k) { return make_double2(A[k], 0.0); });

  double *out = p.sp + (size_t)frame * bins;
  for (int i = tid; i < bins; i += nt) out[i] = exp(W[wb_didx(i)]);
}

}  // namespace

size_t wb_cheaptrick_smem_bytes(int fft_size, int seg_capacity) {
  return sizeof(cplx) * wb_fft_slots(fft_size / 2) + sizeof(double) * ((fft_size + 2) + seg_capacity + 256 + 64);
}

// d_x, d_tpos, d_f0: device.  d_sp: device [f0_length][fft_size/2+1].
// Consumes the global randn stream exactly like the reference's serial loop.
int wb_cheaptrick_run(WbWorkspace *ws, int fs, int fft_size, double q1, double f0_floor_internal,
                      const double *d_x, int x_length, const double *d_tpos, const double *d_f0,
                      int f0_length, double *d_sp, const WbRngCursor &rng, cudaStream_t stream,
                      const WbRowChunks *chunks, const WbFrameRange *range) {
  if (f0_length <= 0) return WB_OK;
  if (range && (range->begin < 0 || range->end > f0_length || range->begin > range->end || chunks)) return WB_ERR_ARG;
  int log2n = 0;
  while ((1 << log2n) < fft_size) ++log2n;
  if ((1 << log2n) != fft_size || fft_size < 128 || fft_size > 16384) return WB_ERR_UNSUPPORTED;
  const int bins = fft_size / 2 + 1;
  const cplx *tw = wb_twiddle_table(fft_size);
  if (!tw) return WB_ERR_CUDA;

  unsigned long long *d_offsets = (unsigned long long *)ws->get("ct_offsets", sizeof(unsigned long long) * (f0_length + 1));
  const int n_rows = range ? range->end - range->begin : f0_length;
  const unsigned long long max_noise = (unsigned long long)(n_rows > 0 ? n_rows : 1) * (unsigned long long)(fft_size + bins);
  double *d_noise = (double *)ws->get("noise_ct", sizeof(double) * max_noise);
  if (!d_offsets || !d_noise) return WB_ERR_CUDA;

  // Row chunks (host API: every finished chunk goes home while the next ones compute) alternate between two streams
  // with a noise buffer each: the randn() fill of a chunk runs beside the frames of the chunk before it.
  const bool row_chunks = chunks && chunks->n > 1 && !range && f0_length >= 64;
  double *d_noise_b = nullptr;
  if (row_chunks) {
    int widest = 0;
    for (int c = 0; c < chunks->n; ++c) widest = wb_max_i(widest, chunks->bounds[c + 1] - chunks->bounds[c]);
    d_noise_b = (double *)ws->get("noise_ct_b", sizeof(double) * (unsigned long long)widest * (unsigned long long)(fft_size + bins));
    if (!d_noise_b) return WB_ERR_CUDA;
  }
  if (rng.wait_skip_in) WB_CUDA_CHECK(cudaStreamWaitEvent(stream, rng.wait_skip_in, 0));
  int rc;
  if (f0_length > WB_SCAN_SINGLE_CTA_MAX) {
    CtCountFn fn = {d_f0, fs, fft_size, f0_floor_internal};
    if ((rc = wb_count_scan_tiles(fn, f0_length, d_offsets, rng.skip_in, nullptr, nullptr, rng.skip_out, ws, "ct_scan_tiles", stream))) return rc;
  } else {
    WB_LAUNCH("ct_count_scan_kernel", ct_count_scan_kernel<<<1, 1024, 0, stream>>>(d_f0, f0_length, fs, fft_size, f0_floor_internal,
                                                                                d_offsets, rng.skip_in, rng.skip_out));  // d_offsets[f0_length] = total
    WB_CUDA_CHECK(cudaGetLastError());
  }
  if (rng.record_skip_out) WB_CUDA_CHECK(cudaEventRecord(rng.record_skip_out, stream));
  const unsigned long long *d_noise_off = d_offsets;
  if (range) {
    // draw only the rows' share of the stream; the frames index it relative to the first row
    unsigned long long *d_rel = (unsigned long long *)ws->get("ct_offsets_rel", sizeof(unsigned long long) * (f0_length + 1));
    unsigned long long *d_pos = (unsigned long long *)ws->get("ct_range_pos", sizeof(unsigned long long) * 2);
    if (!d_rel || !d_pos) return WB_ERR_CUDA;
    if ((rc = wb_range_offsets(d_offsets, *range, d_rel, rng.skip_in, d_pos, d_pos + 1, stream))) return rc;
    if (n_rows > 0 && (rc = wb_rng_fill(rng.state, d_pos, d_pos + 1, max_noise, d_noise, stream))) return rc;
    d_noise_off = d_rel;
  } else if (!row_chunks) {
    rc = wb_rng_fill(rng.state, rng.skip_in, d_offsets + f0_length, max_noise, d_noise, stream);
    if (rc) return rc;
  }

  CtParams p;
  p.x = d_x; p.x_length = x_length; p.tpos = d_tpos; p.f0 = d_f0; p.f0_length = f0_length;
  p.fs = fs; p.fft_size = fft_size; p.log2nc = log2n - 1; p.q1 = q1; p.f0_floor = f0_floor_internal;
  p.twiddle = tw; p.noise = d_noise; p.noise_off = d_noise_off; p.noise_origin = -1;
  p.sp = (range && d_sp) ? d_sp - (size_t)range->begin * bins : d_sp;   // (rows are addressed by absolute frame)
  p.seg_capacity = fft_size / 2 + fft_size / 4 + 8;
  p.error_flag = ws->error_flag();
  const size_t smem = wb_cheaptrick_smem_bytes(fft_size, p.seg_capacity);
  const int threads = wb_max_i(64, wb_min_i(256, fft_size / 8));
  p.frame_begin = range ? range->begin : 0;
  if (n_rows <= 0) {
    // (an empty range still takes part in the stream bookkeeping below)
  } else if (!row_chunks) {
    rc = WB_DISPATCH_LOG2(log2n, 8, 13, {
      if (cudaFuncSetAttribute(ct_frame_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WB_ERR_CUDA;
      WB_LAUNCH("ct_frame_kernel", ct_frame_kernel<L2><<<n_rows, threads, smem, stream>>>(p));
    });
    if (rc) return rc;
    if (chunks) for (int c = 0; c < chunks->n; ++c) WB_CUDA_CHECK(cudaEventRecord(chunks->ev[c], stream));
  } else {
    // everything before this point is on `stream`
    WB_CUDA_CHECK(cudaEventRecord(chunks->ev_ready, stream));
    WB_CUDA_CHECK(cudaStreamWaitEvent(chunks->alt, chunks->ev_ready, 0));
    for (int c = 0; c < chunks->n; ++c) {
      cudaStream_t cs = (c & 1) ? chunks->alt : stream;
      double *nb = (c & 1) ? d_noise_b : d_noise;
      const int cb = chunks->bounds[c], ce = chunks->bounds[c + 1];
      if ((rc = wb_rng_fill(rng.state, rng.skip_in, d_offsets + ce, (unsigned long long)(ce - cb) * (unsigned long long)(fft_size + bins), nb, cs,
                            d_offsets + cb, d_offsets + cb)))
        return rc;
      if (ce > cb) {
        p.frame_begin = cb;
        p.noise = nb;
        p.noise_origin = cb;
        rc = WB_DISPATCH_LOG2(log2n, 8, 13, {
          if (cudaFuncSetAttribute(ct_frame_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WB_ERR_CUDA;
          WbLaunchScope scope("ct_frame_kernel", cs);
          ct_frame_kernel<L2><<<ce - cb, threads, smem, cs>>>(p);
        });
        if (rc) return rc;
      }
      WB_CUDA_CHECK(cudaEventRecord(chunks->ev[c], cs));
    }
    // the caller's stream continues (randn bookkeeping, later calls) after every range
    for (int c = 1; c < chunks->n; c += 2) WB_CUDA_CHECK(cudaStreamWaitEvent(stream, chunks->ev[c], 0));
  }
  WB_CUDA_CHECK(cudaGetLastError());
  return rng.advance ? wb_rng_advance(rng.state, d_offsets + f0_length, rng.skip_in, stream) : WB_OK;
}
