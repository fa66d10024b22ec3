// D4C band aperiodicity with the "Love Train" voiced/unvoiced detector.
// One thread block per frame; every FFT, smoothing pass and the order statistic of the
// band power spectrum stay in shared memory.
//
// Reference: /root/reference/src/d4c.cpp
//   prepareForD4c :60-111, compute :113-173, loveTrain/loveTrainSub :181-240,
//   getWindowedWaveform :246-303, generalBody :308-333, getStaticCentroid :339-360,
//   getCentroid :366-405, getSmoothedPowerSpectrum :411-434, getStaticGroupDelay :440-460,
//   getCoarseAperiodicity :466-503.
#include "wb_internal.h"
#include "wb_fft.cuh"
#include "wb_scan.cuh"
#include "wb_smooth.cuh"

#include <math.h>
#include <vector>

namespace {

#define D4C_HANNING 1
#define D4C_BLACKMAN 2
#define D4C_MAX_AP 8

__device__ __forceinline__ int d4c_half_window(double ratio, int fs, double f0) {
  return wb_round(ratio * fs / f0 / 2.0);  // d4c.cpp:250
}

// ---- randn() call counts and the frames' positions in the stream -------------------------------
__global__ void __launch_bounds__(1024) lt_count_scan_kernel(const double *__restrict__ f0, int n, int fs,
                                                             double lowest_f0, unsigned long long *__restrict__ offsets,
                                                             const unsigned long long *__restrict__ skip_in,
                                                             unsigned long long *__restrict__ skip_out) {
  wb_block_count_scan([&](int i) {
    unsigned long long c = 0;
    if (f0[i] != 0.0) {
      const double cf0 = f0[i] > lowest_f0 ? f0[i] : lowest_f0;
      c = 2ull * d4c_half_window(3.0, fs, cf0) + 1ull;
    }
    return c;
  }, n, offsets, skip_in, skip_out);
}

// The stream position of the body draws is (position before D4C) + (Love Train draws): the two are
// added here, so the Love Train count itself does not have to wait for the stage before D4C.
__global__ void __launch_bounds__(1024) body_count_scan_kernel(const double *__restrict__ f0, const double *__restrict__ ap0,
                                                               int n, int fs, double threshold,
                                                               unsigned long long *__restrict__ offsets,
                                                               const unsigned long long *__restrict__ skip_before,
                                                               const unsigned long long *__restrict__ lt_total,
                                                               unsigned long long *__restrict__ skip_mid,
                                                               unsigned long long *__restrict__ skip_out) {
  if (threadIdx.x == 0) *skip_mid = (skip_before ? *skip_before : 0ull) + *lt_total;
  __syncthreads();
  const unsigned long long *skip_in = skip_mid;
  wb_block_count_scan([&](int i) {
    unsigned long long c = 0;
    if (!(f0[i] == 0 || ap0[i] <= threshold)) {
      const double cf0 = f0[i] > WB_FLOOR_F0_D4C ? f0[i] : WB_FLOOR_F0_D4C;
      c = 3ull * (2ull * d4c_half_window(4.0, fs, cf0) + 1ull);
    }
    return c;
  }, n, offsets, skip_in, skip_out);
}

// the same counts as functors, for the grid-wide scan of long streams (wb_scan.cuh)
struct LtCountFn {
  const double *f0; int fs; double lowest_f0;
  __device__ unsigned long long operator()(int i) const {
    if (f0[i] == 0.0) return 0ull;
    const double cf0 = f0[i] > lowest_f0 ? f0[i] : lowest_f0;
    return 2ull * d4c_half_window(3.0, fs, cf0) + 1ull;
  }
};
struct BodyCountFn {
  const double *f0; const double *ap0; int fs; double threshold;
  __device__ unsigned long long operator()(int i) const {
    if (f0[i] == 0 || ap0[i] <= threshold) return 0ull;
    const double cf0 = f0[i] > WB_FLOOR_F0_D4C ? f0[i] : WB_FLOOR_F0_D4C;
    return 3ull * (2ull * d4c_half_window(4.0, fs, cf0) + 1ull);
  }
};

// ---- Love Train (d4c.cpp:181-240) ---------------------------------------------------------
struct LtParams {
  const double *x; int x_length;
  const double *tpos; const double *f0; int f0_length;
  int fs; int fft_size; int log2nc; double lowest_f0;
  int boundary0, boundary1, boundary2;
  const cplx *twiddle;
  const double *noise; const unsigned long long *noise_off;
  int noise_origin;   // >= 0: the noise buffer starts at the draws of this frame (offsets are taken relative to it)
  double *ap0;
  int frame_begin;   // this launch covers frames frame_begin + blockIdx.x
};

// ---- block-wide reductions with ONE barrier each ----------------------------------------------
// Two scratch buffers alternate (`par`): a buffer is rewritten only after the barrier of the reduction in
// between, which every thread reaches after its reads of that buffer.  red: 2 * 16 * K doubles.  The
// partials are added in warp order, so the result does not depend on scheduling.  All threads call.
template <int K>
__device__ __forceinline__ void d4c_block_sum(double (&v)[K], double *red, int &par) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int j = 0; j < K; ++j) v[j] = wb_warp_sum(v[j]);
  double *buf = red + par * (16 * K);
  par ^= 1;
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < K; ++j) buf[warp * K + j] = v[j];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < K; ++j) v[j] = buf[j];
  for (int w = 1; w < nw; ++w) {
#pragma unroll
    for (int j = 0; j < K; ++j) v[j] += buf[w * K + j];
  }
}
// two sums and two maxima (non-negative values)
__device__ __forceinline__ void d4c_block_sum2_max2(double &sa, double &sb, double &ma, double &mb, double *red, int &par) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  sa = wb_warp_sum(sa);
  sb = wb_warp_sum(sb);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ma = fmax(ma, __shfl_xor_sync(0xffffffffu, ma, o));
    mb = fmax(mb, __shfl_xor_sync(0xffffffffu, mb, o));
  }
  double *buf = red + par * 64;
  par ^= 1;
  if (lane == 0) { buf[warp * 4 + 0] = sa; buf[warp * 4 + 1] = sb; buf[warp * 4 + 2] = ma; buf[warp * 4 + 3] = mb; }
  __syncthreads();
  sa = buf[0]; sb = buf[1]; ma = buf[2]; mb = buf[3];
  for (int w = 1; w < nw; ++w) {
    sa += buf[w * 4 + 0]; sb += buf[w * 4 + 1];
    ma = fmax(ma, buf[w * 4 + 2]); mb = fmax(mb, buf[w * 4 + 3]);
  }
}

// ---- Love Train frames (d4c.cpp:181-240) -------------------------------------------------------------------
// One CTA of T = N/16 threads per frame: the packed real transform (N/2 complex points) has exactly one radix-8
// butterfly per thread and pass, and thread t owns packed samples t + T q, i.e. waveform samples 2 (t + T q) and the
// one after it -- the Blackman window (three periods), the mean removal and the first FFT pass work on registers;
// shared memory only holds the FFT slots (N = 4096: 38 KB, four CTAs per SM at 64 registers).
template <int LOG2N>
__global__ void __launch_bounds__((1 << LOG2N) / 16, ((16384 >> LOG2N) > 16 ? 16 : ((16384 >> LOG2N) < 1 ? 1 : (16384 >> LOG2N))))
lt_frame_kernel(LtParams p) {
  extern __shared__ double2 smem_raw[];
  constexpr int N = 1 << LOG2N, NC = N / 2, T = N / 16;
  cplx *S = smem_raw;
  double *red = reinterpret_cast<double *>(S + wb_fft_slots(NC));   // 2 x 16 x 2
  const int frame = p.frame_begin + blockIdx.x;
  const int tid = threadIdx.x;
  const double f0 = p.f0[frame];
  if (f0 == 0.0) {
    if (tid == 0) p.ap0[frame] = 0.0;
    return;
  }
  const int fs = p.fs;
  const double cf0 = f0 > p.lowest_f0 ? f0 : p.lowest_f0;
  const int hw = d4c_half_window(3.0, fs, cf0);
  const int wlen = 2 * hw + 1;
  const int origin = wb_round(p.tpos[frame] * fs + 0.001);
  const double *noise = p.noise + (p.noise_off[frame] - (p.noise_origin >= 0 ? p.noise_off[p.noise_origin] : 0ull));
  int par = 0;
  double v[16];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = 2 * (tid + q * T) + e;
      v[2 * q + e] = 0.0;
      if (j < wlen) v[2 * q + e] = p.x[wb_min_i(p.x_length - 1, wb_max_i(0, origin + j - hw))];
    }
  }
  // window angle of sample j: pi f0 (2 / 3 / fs) (j - hw) (d4c.cpp:246-283): one sincos at the thread's first sample
  // (the reference's expression), a rotation by one sample for the odd one and by 2 T samples per step (<= 7 steps).
  // The window values are needed twice (windowing, then the mean removal): the recurrence is run twice instead of
  // keeping sixteen more doubles in registers.
  const double wc1 = 2.0 / 3.0 / fs, wc2 = WB_PI * cf0;
  double sd, cd, s1c, c1c, sn0, cs0;
  sincos(wc2 * wc1 * (2 * T), &sd, &cd);
  sincos(wc2 * wc1, &s1c, &c1c);
  sincos(wc2 * (wc1 * (2 * tid - hw)), &sn0, &cs0);
  double s[2] = {0.0, 0.0};
  {
    double sn = sn0, cs = cs0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int j = 2 * (tid + q * T);
      if (j < wlen) {
        const double w0 = 0.42 + 0.5 * cs + 0.08 * (2.0 * cs * cs - 1.0);
        const double v0 = v[2 * q] * w0 + noise[j] * WB_SAFEGUARD;
        v[2 * q] = v0;
        s[0] += v0; s[1] += w0;
        if (j + 1 < wlen) {
          const double c1 = cs * c1c - sn * s1c;
          const double w1 = 0.42 + 0.5 * c1 + 0.08 * (2.0 * c1 * c1 - 1.0);
          const double v1 = v[2 * q + 1] * w1 + noise[j + 1] * WB_SAFEGUARD;
          v[2 * q + 1] = v1;
          s[0] += v1; s[1] += w1;
        }
        const double c_next = cs * cd - sn * sd;
        sn = sn * cd + cs * sd;
        cs = c_next;
      }
    }
  }
  d4c_block_sum<2>(s, red, par);
  const double coef = s[0] / s[1];
  {
    double sn = sn0, cs = cs0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int j = 2 * (tid + q * T);
      if (j < wlen) {
        v[2 * q] -= (0.42 + 0.5 * cs + 0.08 * (2.0 * cs * cs - 1.0)) * coef;
        if (j + 1 < wlen) {
          const double c1 = cs * c1c - sn * s1c;
          v[2 * q + 1] -= (0.42 + 0.5 * c1 + 0.08 * (2.0 * c1 * c1 - 1.0)) * coef;
        }
        const double c_next = cs * cd - sn * sd;
        sn = sn * cd + cs * sd;
        cs = c_next;
      }
    }
  }
  wb_pass_dif8<1, NC, NC>(S, p.twiddle, [&](int, int q) { return make_double2(v[2 * q], v[2 * q + 1]); });
  __syncthreads();
  WbDifPasses<1, NC, NC / 8, 8, true>::run(S, p.twiddle, WbFromSlots());
  // power in (boundary0, boundary1] and (boundary0, boundary2]; bins above N/2 count as zero
  double pw[2] = {0.0, 0.0};
  const int b0 = p.boundary0, b1 = p.boundary1, b2 = p.boundary2;
  auto take = [&](int k, cplx X) {
    if (k > b0 && k <= b2) {
      const double e2 = X.x * X.x + X.y * X.y;
      pw[1] += e2;
      if (k <= b1) pw[0] += e2;
    }
  };
  for (int k = tid; k <= (NC >> 1); k += T) {   // split step of the real transform (see wb_rfft_t)
    if (k == 0) {
      const cplx z = S[0];
      take(0, make_double2(z.x + z.y, 0.0));
      take(NC, make_double2(z.x - z.y, 0.0));
    } else {
      const cplx zk = S[wb_sidx(wb_brev(k, LOG2N - 1))];
      const cplx zc = S[wb_sidx(wb_brev(NC - k, LOG2N - 1))];
      const cplx E = make_double2(0.5 * (zk.x + zc.x), 0.5 * (zk.y - zc.y));
      const cplx O = make_double2(0.5 * (zk.y + zc.y), -0.5 * (zk.x - zc.x));
      const cplx t = wb_cmul(wb_tw<1>(p.twiddle, k), O);
      take(k, wb_cadd(E, t));
      if (k != NC - k) take(NC - k, wb_conj(wb_csub(E, t)));
    }
  }
  d4c_block_sum<2>(pw, red, par);
  if (tid == 0) p.ap0[frame] = pw[0] / pw[1];
}

// ---- linear smoothing (world_common.cpp:82-116, :27-52; interp1Q world_matlabfunctions.cpp:220-241) ----
// The caller has stored P[k], k = 0..NC, at seg[b + k] (b = the boundary of this width); the smoothed values
// of the thread's own bins k = tid + T i come back in out[].  The two interp1Q positions of bin k are
// k + c_lo and k + c_hi with constants c (the knots are the bin grid shifted by half a bin), so the knot index
// and the fraction are formed once instead of per bin through two divisions and a truncation; the reference's
// per-bin expressions give the same numbers up to rounding of the fraction, and the interpolant is
// continuous across a knot, so the results agree to an ulp of the cumulative sum.  All threads call; starts
// with a barrier (P complete), ends WITHOUT one (seg is still being read).
template <int T_, int NC_>
__device__ __forceinline__ void d4c_smooth(double *seg, int b, double width, int fs, double (&out)[9], double *wt) {
  constexpr int N_ = 2 * NC_;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int len = NC_ + 2 * b + 1;
  __syncthreads();
  for (int i = tid; i < b; i += T_) {       // mirrored ends (world_common.cpp:32-47)
    seg[i] = seg[2 * b - i];
    seg[NC_ + b + 1 + i] = seg[NC_ + b - 1 - i];
  }
  __syncthreads();
  // cumulative sum of P * fs / fft_size: sequential inside a thread's chunk, chunk offsets by a warp scan
  const double inv_fft = 1.0 / N_;  // power of two: the product equals the reference's division
  const int chunk = (len + T_ - 1) / T_;
  const int bgn = wb_min_i(len, tid * chunk), end = wb_min_i(len, bgn + chunk);
  double s = 0.0;
  for (int i = bgn; i < end; ++i) { s += seg[i] * fs * inv_fft; seg[i] = s; }
  double incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wt[warp] = incl;
  __syncthreads();
  double off = incl - s;
  for (int w = 0; w < warp; ++w) off += wt[w];
  for (int i = bgn; i < end; ++i) seg[i] += off;
  __syncthreads();
  const double origin = -(b - 0.5) * fs / N_;
  const double inv_interval = static_cast<double>(N_) / fs;
  const double q_lo = (-width / 2.0 - origin) * inv_interval, q_hi = (-width / 2.0 + width - origin) * inv_interval;
  const int b_lo = static_cast<int>(q_lo), b_hi = static_cast<int>(q_hi);
  const double f_lo = q_lo - b_lo, f_hi = q_hi - b_hi;
  const double inv_width = 1.0 / width;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int k = tid + T_ * i;
    out[i] = 0.0;
    if (i < 8 || tid == 0) {
      const double *lo = seg + k + b_lo, *hi = seg + k + b_hi;
      const double low = lo[0] + (lo[1] - lo[0]) * f_lo;
      const double high = hi[0] + (hi[1] - hi[0]) * f_hi;
      out[i] = (high - low) * inv_width;
    }
  }
}

// ---- body (d4c.cpp:308-503 + :155-168) -------------------------------------------------------
struct BodyParams {
  const double *x; int x_length;
  const double *tpos; const double *f0; const double *ap0; int f0_length;
  int fs; double threshold;
  int n_ap; int window_length;         // Nuttall window of the band analysis
  const double *nuttall;               // device, window_length doubles
  const cplx *tw_n;                    // fft_size_d4c entries (real transforms)
  const cplx *tw_2n;                   // 2*fft_size_d4c entries (complex transform of size N)
  const double *noise; const unsigned long long *noise_off;
  int noise_origin;                    // >= 0: the noise buffer starts at the draws of this frame (offsets relative to it)
  int out_fft_size;                    // bins of the output rows
  double *ap;                          // [f0_length][out_fft_size/2+1]
  int *error_flag;
  int frame_begin;                     // this launch covers frames frame_begin + blockIdx.x
};

#define D4C_SPLIT 2      /* row chunks of a split stage (WbStageSplit); measured: 2 -> 1.600 ms per step, 4 -> 1.618 (1.620 unsplit) */
#define D4C_HIST_BINS 256
#define D4C_SEL_WARP_LIST 32

// Shared-memory footprint of the body kernel in bytes (N = D4C FFT size)
static size_t d4c_body_smem_bytes(int N) {
  const int binsp = ((N / 2 + 1) + 1) & ~1;
  return sizeof(cplx) * wb_fft_slots(N) + sizeof(double) * (binsp + 2 * 16 * 5 + 32 + D4C_MAX_AP + 2 + 16) +
         sizeof(int) * (2 * D4C_HIST_BINS + 16);
}

// One CTA of T = N/16 threads per analysed frame: every radix-16 pass of an N-point transform is exactly one
// butterfly per thread, and thread t owns spectrum bins k = t + T i (i < 8; thread 0 also bin N/2) in every
// elementwise step, so spectra travel between the phases in registers.  Shared memory (N = 4096: 95 KB, two
// CTAs per SM):
//   S    : slots of an N-point complex FFT.  While only a packed real transform lives in its lower half, the
//          upper half (segA) is the cumulative-sum array of a linear smoothing; the lower half (segB) serves
//          the next smoothing; the order-statistic lists of the band analysis also live in segA.
//   SC   : static centroid, later the static group delay (random access by the band slices)
template <int LOG2N>
__global__ void __launch_bounds__((1 << LOG2N) / 16, ((8192 >> LOG2N) > 16 ? 16 : ((8192 >> LOG2N) < 1 ? 1 : (8192 >> LOG2N))))
d4c_body_kernel(BodyParams p) {
  extern __shared__ double2 smem_raw[];
  constexpr int N = 1 << LOG2N, NC = N / 2, bins = NC + 1, T = N / 16;
  constexpr int binsp = (bins + 1) & ~1;  // keep 16-byte alignment
  cplx *S = smem_raw;
  double *SC = reinterpret_cast<double *>(S + wb_fft_slots(N));
  double *red = SC + binsp;                      // 2 x 16 x 5
  double *wt = red + 2 * 16 * 5;                 // 32: warp totals of the scans
  double *coarse = wt + 32;                      // D4C_MAX_AP + 2 (+ 16 spare: band power ratios)
  double *ratio = coarse + D4C_MAX_AP + 2;       // D4C_MAX_AP (+ padding)
  int *hist = reinterpret_cast<int *>(ratio + 16);  // 2 x D4C_HIST_BINS counters
  int *ctl = hist + 2 * D4C_HIST_BINS;           // [w]: bucket, [2+w]: rank inside, [4+w]: bucket count, [6+w]: list fill
  double *W = reinterpret_cast<double *>(S);
  double *segA = reinterpret_cast<double *>(S + wb_fft_slots(NC) + 8);
  constexpr int seg_cap = 2 * (wb_fft_slots(N) - wb_fft_slots(NC) - 8);   // doubles in segA
  double *segB = W;                                                       // 2 * slots(NC) >= seg_cap doubles
  (void)W;

  const int frame = p.frame_begin + blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const double f0_in = p.f0[frame];
  if (f0_in == 0 || p.ap0[frame] <= p.threshold) {
    // not analysed: the row keeps its initial value 1 - kMySafeGuardMinimum (d4c.cpp:127-132, :146)
    const int out_bins = p.out_fft_size / 2 + 1;
    double *out = p.ap + (size_t)frame * out_bins;
    for (int i = tid; i < out_bins; i += T) out[i] = 1.0 - WB_SAFEGUARD;
    return;
  }
  const double f0 = f0_in > WB_FLOOR_F0_D4C ? f0_in : WB_FLOOR_F0_D4C;
  const int fs = p.fs;
  const double pos = p.tpos[frame];
  const double *noise = p.noise + (p.noise_off[frame] - (p.noise_origin >= 0 ? p.noise_off[p.noise_origin] : 0ull));
  int par = 0;

  const int hw = d4c_half_window(4.0, fs, f0);
  const int wlen = 2 * hw + 1;
  const int b_sp = static_cast<int>(f0 * N / fs) + 1;          // smoothing boundaries (world_common.cpp:88)
  const int b_g1 = static_cast<int>(f0 / 2.0 * N / fs) + 1;
  if (NC + 2 * b_sp + 1 > seg_cap || 2 + static_cast<int>(f0 * N / fs) > 4 * T) {
    if (tid == 0) atomicExch(p.error_flag, WB_ERR_UNSUPPORTED);
    return;
  }
  // the noise of the later windows is on its way to L2 while the first window is built
  {
    const char *nb = reinterpret_cast<const char *>(noise + wlen);
    const int lines = (2 * wlen * 8 + 127) / 128 + 1;
    for (int i = tid; i < lines; i += T) asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + (size_t)i * 128));
  }

  // ---- static centroid: two windows at pos -/+ 0.25/f0 (d4c.cpp:339-405)
  // The reference transforms w[n] and (n+1) w[n] separately; both are real, so one complex transform of
  // z[n] = w[n] + i (n+1) w[n] yields both spectra.  Thread t owns samples t + T q -- exactly one radix-16
  // butterfly of the first FFT pass -- so the windowed, mean-free, unit-power waveform (d4c.cpp:246-303,
  // :372-380) is built in registers and handed to the transform without touching shared memory.
  // The window angle of sample j is linear in j: one sincos at the thread's first sample (the reference's
  // expression), then a rotation by T samples per step (<= 15 steps, ~1e-15 drift); cos(2t) = 2 cos(t)^2 - 1
  // for the Blackman term (d4c.cpp:266-283).  The angles are the same in all three windows.
  const double wc1 = 2.0 / 4.0 / fs, wc2 = WB_PI * f0;
  double rot_sd, rot_cd, rot_sn, rot_cs;
  sincos(wc2 * wc1 * T, &rot_sd, &rot_cd);
  sincos(wc2 * (wc1 * (tid - hw)), &rot_sn, &rot_cs);
  for (int c = 0; c < 2; ++c) {
    const double cpos = (c == 0) ? pos - 0.25 / f0 : pos + 0.25 / f0;
    const int origin = wb_round(cpos * fs + 0.001);
    double v[16], w[16];
    // all waveform / noise loads first (independent: they overlap), parked in v[] / w[] until they are used
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int j = tid + q * T;
      v[q] = 0.0; w[q] = 0.0;
      if (j < wlen) {
        v[q] = p.x[wb_min_i(p.x_length - 1, wb_max_i(0, origin + j - hw))];
        w[q] = noise[j];
      }
    }
    double sn = rot_sn, cs = rot_cs;
    double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};   // sum v, sum w, sum v^2, sum v w, sum w^2
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int j = tid + q * T;
      if (j < wlen) {
        const double wq = 0.42 + 0.5 * cs + 0.08 * (2.0 * cs * cs - 1.0);
        const double vq = v[q] * wq + w[q] * WB_SAFEGUARD;
        v[q] = vq; w[q] = wq;
        s[0] += vq; s[1] += wq; s[2] += vq * vq; s[3] += vq * wq; s[4] += wq * wq;
        const double c_next = cs * rot_cd - sn * rot_sd;
        sn = sn * rot_cd + cs * rot_sd;
        cs = c_next;
      }
    }
    // One reduction gives the mean-removal coefficient AND the power of the mean-free waveform
    // (sum (v - c w)^2 = sum v^2 - 2 c sum v w + c^2 sum w^2), so the unit-power scale (d4c.cpp:372-380) needs no
    // second pass over the block; one reciprocal per window instead of a division per sample (quotients
    // within an ulp of the reference's).
    d4c_block_sum<5>(s, red, par);
    const double coef = s[0] / s[1];
    const double inv_power = 1.0 / sqrt(s[2] - 2.0 * coef * s[3] + coef * coef * s[4]);
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = (v[q] - w[q] * coef) * inv_power;   // (samples past the window stay 0)
    noise += wlen;
    wb_cfft_dif_in_t<1, LOG2N, 16, true>(S, p.tw_2n, [&](int j, int q) { return make_double2(v[q], v[q] * (j + 1.0)); });
    // spectra of w (X1) and (n+1) w (X2) from Z[k], Z[N-k]; centroid term Re X1 Re X2 + Im X1 Im X2 (d4c.cpp:400)
    // = (Re Z[k] Im Z[N-k] + Im Z[k] Re Z[N-k]) / 2
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const int k = tid + T * i;
      if (i < 8 || tid == 0) {
        const cplx zk = S[wb_sidx(wb_brev(k, LOG2N))];
        const cplx zc = S[wb_sidx(wb_brev((N - k) & (N - 1), LOG2N))];
        const double cen = 0.5 * (zk.x * zc.y + zk.y * zc.x);
        SC[k] = (c == 0) ? cen : SC[k] + cen;
      }
    }
  }

  // ---- smoothed power spectrum (d4c.cpp:411-434): Hanning window at pos, real transform as a packed N/2-point
  // complex one whose first pass is radix 8 (one butterfly per thread: packed samples t + T q, i.e. waveform
  // samples 2 (t + T q) and the one after it, again straight from registers)
  {
    const int origin = wb_round(pos * fs + 0.001);
    double v[16], w[16];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 2 * (tid + q * T) + e;
        v[2 * q + e] = 0.0; w[2 * q + e] = 0.0;
        if (j < wlen) {
          v[2 * q + e] = p.x[wb_min_i(p.x_length - 1, wb_max_i(0, origin + j - hw))];
          w[2 * q + e] = noise[j];
        }
      }
    }
    // angles: sample 2 t first, one sample further for the odd one, 2 T samples per step (double angle of the
    // T-sample rotation)
    double s1c, c1c, sn, cs;
    sincos(wc2 * wc1, &s1c, &c1c);
    sincos(wc2 * (wc1 * (2 * tid - hw)), &sn, &cs);
    const double sd2 = 2.0 * rot_sd * rot_cd, cd2 = 2.0 * rot_cd * rot_cd - 1.0;
    double s[2] = {0.0, 0.0};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int j = 2 * (tid + q * T);
      if (j < wlen) {
        const double w0 = 0.5 * cs + 0.5;
        const double v0 = v[2 * q] * w0 + w[2 * q] * WB_SAFEGUARD;
        v[2 * q] = v0; w[2 * q] = w0;
        s[0] += v0; s[1] += w0;
        if (j + 1 < wlen) {
          const double w1 = 0.5 * (cs * c1c - sn * s1c) + 0.5;
          const double v1 = v[2 * q + 1] * w1 + w[2 * q + 1] * WB_SAFEGUARD;
          v[2 * q + 1] = v1; w[2 * q + 1] = w1;
          s[0] += v1; s[1] += w1;
        }
        const double c_next = cs * cd2 - sn * sd2;
        sn = sn * cd2 + cs * sd2;
        cs = c_next;
      }
    }
    d4c_block_sum<2>(s, red, par);
    const double coef = s[0] / s[1];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] -= w[q] * coef;
    wb_pass_dif8<1, NC, NC>(S, p.tw_n, [&](int, int q) { return make_double2(v[2 * q], v[2 * q + 1]); });
    __syncthreads();
    WbDifPasses<1, NC, NC / 8, 16, true>::run(S, p.tw_n, WbFromSlots());
    // split step of the real transform (see wb_rfft_t): power of bins k and NC - k, stored for the smoothing
    double *P = segA + b_sp;
    for (int k = tid; k <= (NC >> 1); k += T) {
      if (k == 0) {
        const cplx z = S[0];
        P[0] = (z.x + z.y) * (z.x + z.y);
        P[NC] = (z.x - z.y) * (z.x - z.y);
      } else {
        const cplx zk = S[wb_sidx(wb_brev(k, LOG2N - 1))];
        const cplx zc = S[wb_sidx(wb_brev(NC - k, LOG2N - 1))];
        const cplx E = make_double2(0.5 * (zk.x + zc.x), 0.5 * (zk.y - zc.y));
        const cplx O = make_double2(0.5 * (zk.y + zc.y), -0.5 * (zk.x - zc.x));
        const cplx t = wb_cmul(wb_tw<1>(p.tw_n, k), O);
        const cplx X = wb_cadd(E, t), Y = wb_csub(E, t);
        P[k] = X.x * X.x + X.y * X.y;
        if (k != NC - k) P[NC - k] = Y.x * Y.x + Y.y * Y.y;
      }
    }
    __syncthreads();
    // DC correction of both spectra (world_common.cpp:61-80): bins below f0 receive the interpolated replica
    // P(f0 - f).  (xi - f0) / (-fs / N) is formed as a product with N / fs: the quotient can differ in the last
    // bit, the interpolant is continuous.)
    {
      const int upper_limit = 2 + static_cast<int>(f0 * N / fs);
      const int n_rep = upper_limit - 1;
      const double inv_dx = -static_cast<double>(N) / fs;
      double rep_c[4], rep_p[4];
      int cnt = 0;
      for (int i = tid; i < n_rep && cnt < 4; i += T, ++cnt) {
        const double xi = static_cast<double>(i) * fs / N;
        const double qd = (xi - f0) * inv_dx;
        const int base = static_cast<int>(qd);
        const double frac = qd - base;
        rep_c[cnt] = SC[base] + (SC[base + 1] - SC[base]) * frac;
        rep_p[cnt] = P[base] + (P[base + 1] - P[base]) * frac;
      }
      __syncthreads();
      cnt = 0;
      for (int i = tid; i < n_rep && cnt < 4; i += T, ++cnt) { SC[i] += rep_c[cnt]; P[i] += rep_p[cnt]; }
    }
  }
  double sgd[9], t1[9];
  d4c_smooth<T, NC>(segA, b_sp, f0, fs, t1, wt);

  // ---- static group delay (d4c.cpp:440-460): centroid / smoothed power, smoothed with f0 / 2, minus its own
  // smoothing with f0
  {
    double *P = segB + b_g1;
#pragma unroll
    for (int i = 0; i < 9; ++i)
      if (i < 8 || tid == 0) P[tid + T * i] = SC[tid + T * i] / t1[i];
  }
  d4c_smooth<T, NC>(segB, b_g1, f0 / 2.0, fs, sgd, wt);
  {
    double *P = segA + b_sp;
#pragma unroll
    for (int i = 0; i < 9; ++i)
      if (i < 8 || tid == 0) P[tid + T * i] = sgd[i];
  }
  d4c_smooth<T, NC>(segA, b_sp, f0, fs, t1, wt);
#pragma unroll
  for (int i = 0; i < 9; ++i)
    if (i < 8 || tid == 0) SC[tid + T * i] = sgd[i] - t1[i];
  // (the barriers inside the first band transform order these writes before the slices are read: the slices are
  // read in its first pass -- so one barrier here)
  __syncthreads();

  // ---- coarse aperiodicity per 3 kHz band (d4c.cpp:466-503)
  const int wl = p.window_length, hwl = wl / 2;
  const int boundary = wb_round(N * 8.0 / wl);
  const int n_top = boundary + 1;   // the reference sorts the band's power spectrum and sums the bins - boundary - 1
                                    // smallest values (cumulative-sum index bins - boundary - 2): all but the n_top largest
  // Bands are transformed in pairs: band b in the real part, band b+1 in the imaginary part of ONE complex
  // N-point transform (both are real, so the spectra separate exactly).  The power values stay in registers;
  // the order statistic is a top-n_top selection: a histogram over the distance from the maximum in units of
  // 2^49 key steps (1/8 of a binade), counted from the top, gives the bucket holding the n_top-th largest value;
  // everything in lower buckets is summed directly, the bucket itself is ranked (<= 32 keys: one warp;
  // otherwise a block-wide bitonic sort of the bucket).
  unsigned long long *list = reinterpret_cast<unsigned long long *>(segA);   // 2 x bins keys
  for (int b = 0; b < p.n_ap; b += 2) {
    const bool has2 = (b + 1 < p.n_ap);
    const int center_a = static_cast<int>(WB_FREQ_INTERVAL * (b + 1) * N / fs);
    const int center_b = static_cast<int>(WB_FREQ_INTERVAL * (b + 2) * N / fs);
    for (int i = tid; i < 2 * D4C_HIST_BINS; i += T) hist[i] = 0;
    if (tid < 2) ctl[6 + tid] = 0;
    // the windowed band slices feed the first FFT pass directly (the rest of the N points is zero padding)
    wb_cfft_dif_in_t<1, LOG2N, 16, true>(S, p.tw_2n, [&](int j, int) {
      cplx z = make_double2(0.0, 0.0);
      if (j < wl) {
        const double nw = __ldg(&p.nuttall[j]);
        z.x = SC[center_a - hwl + j] * nw;
        if (has2) z.y = SC[center_b - hwl + j] * nw;
      }
      return z;
    });
    double pa[9], pb[9];
    double tot_a = 0.0, tot_b = 0.0, max_a = 0.0, max_b = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const int k = tid + T * i;
      pa[i] = 0.0; pb[i] = 0.0;
      if (i < 8 || tid == 0) {
        const cplx zk = S[wb_sidx(wb_brev(k, LOG2N))];
        const cplx zc = S[wb_sidx(wb_brev((N - k) & (N - 1), LOG2N))];
        const double ar = 0.5 * (zk.x + zc.x), ai = 0.5 * (zk.y - zc.y);   // spectrum of band b
        const double br = 0.5 * (zk.y + zc.y), bi = -0.5 * (zk.x - zc.x);  // spectrum of band b+1
        pa[i] = ar * ar + ai * ai;
        pb[i] = br * br + bi * bi;
        tot_a += pa[i]; tot_b += pb[i];
        max_a = fmax(max_a, pa[i]); max_b = fmax(max_b, pb[i]);
      }
    }
    d4c_block_sum2_max2(tot_a, tot_b, max_a, max_b, red, par);   // (S is free after this barrier)
    const unsigned long long mk_a = (unsigned long long)__double_as_longlong(max_a);
    const unsigned long long mk_b = (unsigned long long)__double_as_longlong(max_b);
    auto digit = [](unsigned long long mk, double v) {
      const unsigned long long d = (mk - (unsigned long long)__double_as_longlong(v)) >> 49;
      return (int)(d < (unsigned long long)(D4C_HIST_BINS - 1) ? d : (unsigned long long)(D4C_HIST_BINS - 1));
    };
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      if (i < 8 || tid == 0) {
        atomicAdd(&hist[digit(mk_a, pa[i])], 1);
        if (has2) atomicAdd(&hist[D4C_HIST_BINS + digit(mk_b, pb[i])], 1);
      }
    }
    __syncthreads();
    // bucket of the n_top-th largest value: exclusive scan of the counters from the top
    {
      constexpr int PER = (D4C_HIST_BINS + T - 1) / T;
      int cnt[2][PER], tot[2] = {0, 0}, incl[2];
#pragma unroll
      for (int w = 0; w < 2; ++w) {
#pragma unroll
        for (int q = 0; q < PER; ++q) {
          const int bin = tid * PER + q;
          cnt[w][q] = bin < D4C_HIST_BINS ? hist[w * D4C_HIST_BINS + bin] : 0;
          tot[w] += cnt[w][q];
        }
        incl[w] = tot[w];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl[w], o);
          if (lane >= o) incl[w] += t;
        }
      }
      int *wtot = reinterpret_cast<int *>(wt);
      if (lane == 31) { wtot[warp] = incl[0]; wtot[16 + warp] = incl[1]; }
      __syncthreads();
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        int excl = incl[w] - tot[w];
        for (int q = 0; q < warp; ++q) excl += wtot[w * 16 + q];
#pragma unroll
        for (int q = 0; q < PER; ++q) {
          if (n_top > excl && n_top <= excl + cnt[w][q]) {
            ctl[w] = tid * PER + q;
            ctl[2 + w] = n_top - excl;      // that many of the bucket's largest keys are left out
            ctl[4 + w] = cnt[w][q];
          }
          excl += cnt[w][q];
        }
      }
      __syncthreads();
    }
    const int bk_a = ctl[0], bk_b = ctl[1];
    double low_a = 0.0, low_b = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      if (i < 8 || tid == 0) {
        const int da = digit(mk_a, pa[i]);
        if (da > bk_a) low_a += pa[i];
        else if (da == bk_a) list[atomicAdd(&ctl[6], 1)] = (unsigned long long)__double_as_longlong(pa[i]);
        if (has2) {
          const int db = digit(mk_b, pb[i]);
          if (db > bk_b) low_b += pb[i];
          else if (db == bk_b) list[bins + atomicAdd(&ctl[7], 1)] = (unsigned long long)__double_as_longlong(pb[i]);
        }
      }
    }
    __syncthreads();
    // the bucket's keys in descending order; those after the first r join the sum, added in rank order (the
    // collection order above depends on the atomics and must not leak into the rounding)
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      if (w == 1 && !has2) break;
      const int n = ctl[4 + w], r = ctl[2 + w];
      unsigned long long *lw = list + w * bins;
      if (n <= D4C_SEL_WARP_LIST) {
        if (warp == (w % (T / 32))) {
          const unsigned long long mine = lane < n ? lw[lane] : 0ull;
          int rank = 0;
          for (int j = 0; j < n; ++j) {
            const unsigned long long other = __shfl_sync(0xffffffffu, mine, j);
            rank += (other > mine || (other == mine && j < lane)) ? 1 : 0;
          }
          __syncwarp();
          if (lane < n) lw[rank] = mine;
          __syncwarp();
          double vsum = (lane < n && lane >= r) ? __longlong_as_double((long long)lw[lane]) : 0.0;
          vsum = wb_warp_sum(vsum);
          if (lane == 0) { if (w == 0) low_a += vsum; else low_b += vsum; }
        }
      } else {
        int pw2 = 1;
        while (pw2 < n) pw2 <<= 1;
        for (int i = n + tid; i < pw2; i += T) lw[i] = 0ull;
        __syncthreads();
        for (int kk = 2; kk <= pw2; kk <<= 1) {
          for (int jj = kk >> 1; jj > 0; jj >>= 1) {
            for (int i = tid; i < pw2; i += T) {
              const int ixj = i ^ jj;
              if (ixj > i) {
                const unsigned long long x0 = lw[i], x1 = lw[ixj];
                const bool desc = (i & kk) == 0;
                if ((x0 < x1) == desc) { lw[i] = x1; lw[ixj] = x0; }
              }
            }
            __syncthreads();
          }
        }
        if (warp == 0) {   // rank order again: one warp adds the tail with a fixed association
          double vsum = 0.0;
          for (int i = r + lane; i < n; i += 32) vsum += __longlong_as_double((long long)lw[i]);
          vsum = wb_warp_sum(vsum);
          if (lane == 0) { if (w == 0) low_a += vsum; else low_b += vsum; }
        }
      }
    }
    double lows[2] = {low_a, low_b};
    d4c_block_sum<2>(lows, red, par);
    if (tid == 0) {
      ratio[b] = lows[0] / tot_a;
      if (has2) ratio[b + 1] = lows[1] / tot_b;
    }
  }
  __syncthreads();
  if (tid == T - 1) {
    coarse[0] = -60.0;                      // d4c.cpp:84
    coarse[p.n_ap + 1] = -WB_SAFEGUARD;     // d4c.cpp:85
  }
  if (tid < p.n_ap) {                       // one band per thread (the logarithms used to run one after the other)
    const double ca = 10 * log10(ratio[tid]) + (f0 - 100) / 50.0;  // d4c.cpp:325-327
    coarse[tid + 1] = ca < 0.0 ? ca : 0.0;
  }
  __syncthreads();

  // ---- interp1 to the output bins + 10^(v/20) (d4c.cpp:160-168)
  const int out_bins = p.out_fft_size / 2 + 1;
  double *out = p.ap + (size_t)frame * out_bins;
  const int nk = p.n_ap + 2;
  for (int i = tid; i < out_bins; i += T) {
    const double xi = static_cast<double>(i) * fs / p.out_fft_size;
    // histc (world_matlabfunctions.cpp:136-156): first knot strictly above xi, clamped to [1, nk-1]
    int k = 1;
    while (k < nk - 1) {
      if (xi < k * WB_FREQ_INTERVAL) break;
      ++k;
    }
    const double x0 = (k - 1) * WB_FREQ_INTERVAL;
    const double x1 = (k == nk - 1) ? fs / 2.0 : k * WB_FREQ_INTERVAL;
    const double s = (xi - x0) / (x1 - x0);
    const double v = coarse[k - 1] + s * (coarse[k] - coarse[k - 1]);
    out[i] = exp(v * 0.11512925464970228420);  // 10^(v / 20) = e^(v ln 10 / 20): within an ulp or two of pow(10, v / 20)
  }
}

int ilog2_exact(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return ((1 << l) == n) ? l : -1;
}

}  // namespace

int wb_d4c_fft_size(int fs) {  // d4c.cpp:63-64
  return (int)pow(2.0, 1.0 + (int)(log(4.0 * fs / WB_FLOOR_F0_D4C + 1) / WB_LOG2));
}
int wb_d4c_lt_fft_size(int fs) {  // d4c.cpp:101-103
  return (int)pow(2.0, 1.0 + (int)(log(3.0 * fs / 40.0 + 1) / WB_LOG2));
}
int wb_number_of_aperiodicities(int fs) {  // d4c.cpp:65-67, codec.cpp:211-214
  const double a = fs / 2.0 - WB_FREQ_INTERVAL;
  return static_cast<int>((WB_UPPER_LIMIT < a ? WB_UPPER_LIMIT : a) / WB_FREQ_INTERVAL);
}

int wb_d4c_run(WbWorkspace *ws, int fs, double threshold, const double *d_x, int x_length, const double *d_tpos,
               const double *d_f0, int f0_length, int out_fft_size, double *d_ap, const WbRngCursor &rng,
               cudaStream_t stream, const WbRowChunks *chunks, const WbFrameRange *range, int phase, double *d_ap0_ext,
               const WbStageSplit *split) {
  if (f0_length <= 0) return WB_OK;
  if (split && (range || chunks || phase != 0 || f0_length < 64)) split = nullptr;   // whole-utterance calls only
  if (range && (range->begin < 0 || range->end > f0_length || range->begin > range->end || chunks)) return WB_ERR_ARG;
  if (phase < 0 || phase > 2 || (phase != 0 && !d_ap0_ext)) return WB_ERR_ARG;
  const int row0 = range ? range->begin : 0;
  const int n_rows = range ? range->end - range->begin : f0_length;
  const int N = wb_d4c_fft_size(fs), N_lt = wb_d4c_lt_fft_size(fs);
  const int l = ilog2_exact(N), l_lt = ilog2_exact(N_lt);
  if (l < 9 || N > 8192 || l_lt < 7 || N_lt > 16384) return WB_ERR_UNSUPPORTED;
  const int n_ap = wb_number_of_aperiodicities(fs);
  if (n_ap < 0 || n_ap > D4C_MAX_AP) return WB_ERR_UNSUPPORTED;
  const int out_bins = out_fft_size / 2 + 1;

  (void)out_bins;
  unsigned long long *d_offsets = (unsigned long long *)ws->get("d4c_offsets", sizeof(unsigned long long) * (f0_length + 1));
  double *d_ap0 = d_ap0_ext ? d_ap0_ext : (double *)ws->get("d4c_ap0", sizeof(double) * f0_length);
  const unsigned long long max_noise_lt = (unsigned long long)(n_rows > 0 ? n_rows : 1) * N_lt;
  const unsigned long long max_noise_body = (unsigned long long)(n_rows > 0 ? n_rows : 1) * 3ull * N;
  // sharded streams: the rows' share of the stream, indexed relative to the first row (see WbFrameRange)
  unsigned long long *d_rel = nullptr, *d_pos = nullptr;
  if (range) {
    d_rel = (unsigned long long *)ws->get("d4c_offsets_rel", sizeof(unsigned long long) * (f0_length + 1));
    d_pos = (unsigned long long *)ws->get("d4c_range_pos", sizeof(unsigned long long) * 2);
    if (!d_rel || !d_pos) return WB_ERR_CUDA;
  }
  double *d_noise = (double *)ws->get("noise_d4c", sizeof(double) * (max_noise_body > max_noise_lt ? max_noise_body : max_noise_lt));
  // stream position after the Love Train draws (= skip_in + Love Train count)
  unsigned long long *d_skip_mid = (unsigned long long *)ws->get("d4c_skip_mid", sizeof(unsigned long long));
  unsigned long long *d_skip_end = rng.skip_out ? rng.skip_out : (unsigned long long *)ws->get("d4c_skip_end", sizeof(unsigned long long));
  if (!d_offsets || !d_ap0 || !d_noise || !d_skip_mid || !d_skip_end) return WB_ERR_CUDA;
  const cplx *tw_lt = wb_twiddle_table(N_lt);
  const cplx *tw_n = wb_twiddle_table(N);
  const cplx *tw_2n = wb_twiddle_table(2 * N);
  if (!tw_lt || !tw_n || !tw_2n) return WB_ERR_CUDA;

  // Nuttall window of the band analysis (d4c.cpp:69-72, world_common.cpp:118-126), host libm like the reference
  const int window_length = static_cast<int>(WB_FREQ_INTERVAL * N / fs) * 2 + 1;
  // (cached per workspace: the upload below is skipped -- and absent from a captured graph -- once the
  // table of this length sits in this buffer)
  double *d_nuttall = (double *)ws->get("d4c_nuttall", sizeof(double) * window_length);
  int *nuttall_tag = (int *)ws->get_pinned("d4c_nuttall_tag", sizeof(int) * 4);
  if (!d_nuttall || !nuttall_tag) return WB_ERR_CUDA;
  const long long tag_ptr = (long long)(size_t)d_nuttall;
  if (nuttall_tag[0] != window_length || nuttall_tag[1] != (int)(tag_ptr & 0x7fffffff) || nuttall_tag[2] != (int)(tag_ptr >> 31)) {
    double *h = (double *)ws->get_pinned("d4c_nuttall_h", sizeof(double) * window_length);
    if (!h) return WB_ERR_CUDA;
    for (int i = 0; i < window_length; ++i) {
      const double tmp = i / (window_length - 1.0);
      h[i] = 0.355768 - 0.487396 * cos(2.0 * WB_PI * tmp) + 0.144232 * cos(4.0 * WB_PI * tmp) -
             0.012604 * cos(6.0 * WB_PI * tmp);
    }
    WB_CUDA_CHECK(cudaMemcpyAsync(d_nuttall, h, sizeof(double) * window_length, cudaMemcpyHostToDevice, stream));
    nuttall_tag[0] = window_length; nuttall_tag[1] = (int)(tag_ptr & 0x7fffffff); nuttall_tag[2] = (int)(tag_ptr >> 31);
  }

  // ---- Love Train (its frame offsets do not depend on the stream position: only the draw itself waits)
  unsigned long long *d_lt_total = (unsigned long long *)ws->get("d4c_lt_total", sizeof(unsigned long long));
  if (!d_lt_total) return WB_ERR_CUDA;
  int rc;
  if (f0_length > WB_SCAN_SINGLE_CTA_MAX) {
    LtCountFn fn = {d_f0, fs, 40.0};
    if ((rc = wb_count_scan_tiles(fn, f0_length, d_offsets, nullptr, nullptr, nullptr, d_lt_total, ws, "d4c_scan_tiles", stream))) return rc;
  } else {
    WB_LAUNCH("lt_count_scan_kernel", lt_count_scan_kernel<<<1, 1024, 0, stream>>>(d_f0, f0_length, fs, 40.0, d_offsets, nullptr, d_lt_total));
    WB_CUDA_CHECK(cudaGetLastError());
  }
  if (rng.wait_skip_in) WB_CUDA_CHECK(cudaStreamWaitEvent(stream, rng.wait_skip_in, 0));
  // Row chunks of a whole-utterance call alternate between two streams, one noise buffer each: the randn() fill of a
  // chunk runs beside the frames of the chunk before it (integer work beside fp64 work) instead of in front of them.
  // Two users: the fused pipeline (WbStageSplit: D4C_SPLIT chunks) and the host API's chunked download (WbRowChunks:
  // an event per chunk for the copy stream).
  const bool row_chunks = chunks && chunks->n > 1 && !range && phase == 0 && f0_length >= 64;
  const int n_chunks = split ? D4C_SPLIT : row_chunks ? chunks->n : 1;
  cudaStream_t alt = split ? split->alt : row_chunks ? chunks->alt : nullptr;
  auto chunk_bound = [&](int c) -> int {
    return split ? (int)((long long)f0_length * c / D4C_SPLIT) : chunks->bounds[c];
  };
  double *d_noise_b = nullptr;
  if (n_chunks > 1) {
    int widest = 0;
    for (int c = 0; c < n_chunks; ++c) widest = max(widest, chunk_bound(c + 1) - chunk_bound(c));
    const unsigned long long need_lt = (unsigned long long)(f0_length / 2 + 2) * N_lt, need_body = (unsigned long long)widest * 3ull * N;
    d_noise_b = (double *)ws->get("noise_d4c_b", sizeof(double) * (need_lt > need_body ? need_lt : need_body));
    if (!d_noise_b) return WB_ERR_CUDA;
  }
  if (phase != 2) {
  LtParams p;
  p.x = d_x; p.x_length = x_length; p.tpos = d_tpos; p.f0 = d_f0; p.f0_length = f0_length;
  p.fs = fs; p.fft_size = N_lt; p.log2nc = l_lt - 1; p.lowest_f0 = 40.0;
  p.boundary0 = static_cast<int>(ceil(100.0 * N_lt / fs));
  p.boundary1 = static_cast<int>(ceil(4000.0 * N_lt / fs));
  p.boundary2 = static_cast<int>(ceil(7900.0 * N_lt / fs));
  p.twiddle = tw_lt; p.ap0 = d_ap0;
  const size_t smem = sizeof(cplx) * wb_fft_slots(N_lt / 2) + sizeof(double) * (2 * 16 * 2);
  auto launch_lt = [&](int begin, int count, const double *noise, const unsigned long long *off, int origin, cudaStream_t cs) -> int {
    if (count <= 0) return WB_OK;
    p.noise = noise; p.noise_off = off; p.noise_origin = origin; p.frame_begin = begin;
    int r2 = WB_DISPATCH_LOG2(l_lt, 9, 14, {
      if (cudaFuncSetAttribute(lt_frame_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WB_ERR_CUDA;
      WbLaunchScope scope("lt_frame_kernel", cs);
      lt_frame_kernel<L2><<<count, N_lt / 16, smem, cs>>>(p);   // one radix-8 butterfly per thread and pass
    });
    if (r2) return r2;
    WB_CUDA_CHECK(cudaGetLastError());
    return WB_OK;
  };
  if (n_chunks > 1) {
    // two halves (the decisions of all rows are needed before the body's draws can be counted)
    cudaEvent_t fork = split ? split->fork[0] : chunks->ev_ready, join = split ? split->join[0] : chunks->ev[0];
    WB_CUDA_CHECK(cudaEventRecord(fork, stream));
    WB_CUDA_CHECK(cudaStreamWaitEvent(alt, fork, 0));
    for (int c = 0; c < 2; ++c) {
      const int cb = (int)((long long)f0_length * c / 2), ce = (int)((long long)f0_length * (c + 1) / 2);
      cudaStream_t cs = (c & 1) ? alt : stream;
      double *nb = (c & 1) ? d_noise_b : d_noise;
      if ((rc = wb_rng_fill(rng.state, rng.skip_in, d_offsets + ce, (unsigned long long)(ce - cb) * N_lt, nb, cs, d_offsets + cb, d_offsets + cb)))
        return rc;
      if ((rc = launch_lt(cb, ce - cb, nb, d_offsets, cb, cs))) return rc;
    }
    WB_CUDA_CHECK(cudaEventRecord(join, alt));
    WB_CUDA_CHECK(cudaStreamWaitEvent(stream, join, 0));
  } else {
    if (range) {
      if ((rc = wb_range_offsets(d_offsets, *range, d_rel, rng.skip_in, d_pos, d_pos + 1, stream))) return rc;
      if (n_rows > 0 && (rc = wb_rng_fill(rng.state, d_pos, d_pos + 1, max_noise_lt, d_noise, stream))) return rc;
    } else if ((rc = wb_rng_fill(rng.state, rng.skip_in, d_offsets + f0_length, max_noise_lt, d_noise, stream))) return rc;
    if ((rc = launch_lt(row0, n_rows, d_noise, range ? d_rel : d_offsets, -1, stream))) return rc;
  }
  }
  if (phase == 1) return WB_OK;   // (the stream bookkeeping is done by the body phase)
  // ---- body
  if (f0_length > WB_SCAN_SINGLE_CTA_MAX) {
    BodyCountFn fn = {d_f0, d_ap0, fs, threshold};
    if ((rc = wb_count_scan_tiles(fn, f0_length, d_offsets, rng.skip_in, d_lt_total, d_skip_mid, d_skip_end, ws, "d4c_scan_tiles", stream))) return rc;
  } else {
    WB_LAUNCH("body_count_scan_kernel", body_count_scan_kernel<<<1, 1024, 0, stream>>>(d_f0, d_ap0, f0_length, fs, threshold, d_offsets,
                                                                                    rng.skip_in, d_lt_total, d_skip_mid, d_skip_end));
    WB_CUDA_CHECK(cudaGetLastError());
  }
  if (rng.record_skip_out) WB_CUDA_CHECK(cudaEventRecord(rng.record_skip_out, stream));
  if (n_chunks > 1) {
    cudaEvent_t fork = split ? split->fork[1] : chunks->ev_ready;
    WB_CUDA_CHECK(cudaEventRecord(fork, stream));
    WB_CUDA_CHECK(cudaStreamWaitEvent(alt, fork, 0));
    // (the fills are enqueued chunk by chunk together with the frame kernels below)
  } else if (range) {
    if ((rc = wb_range_offsets(d_offsets, *range, d_rel, d_skip_mid, d_pos, d_pos + 1, stream))) return rc;
    if (n_rows > 0 && (rc = wb_rng_fill(rng.state, d_pos, d_pos + 1, max_noise_body, d_noise, stream))) return rc;
  } else if ((rc = wb_rng_fill(rng.state, d_skip_mid, d_offsets + f0_length, max_noise_body, d_noise, stream))) return rc;
  if (n_rows > 0) {
    BodyParams p;
    p.x = d_x; p.x_length = x_length; p.tpos = d_tpos; p.f0 = d_f0; p.ap0 = d_ap0; p.f0_length = f0_length;
    p.fs = fs; p.threshold = threshold;
    p.n_ap = n_ap; p.window_length = window_length; p.nuttall = d_nuttall;
    p.tw_n = tw_n; p.tw_2n = tw_2n; p.noise = d_noise; p.noise_off = range ? d_rel : d_offsets; p.noise_origin = -1;
    p.out_fft_size = out_fft_size;
    p.ap = range ? d_ap - (size_t)row0 * (out_fft_size / 2 + 1) : d_ap;   // (rows are addressed by absolute frame)
    p.error_flag = ws->error_flag();
    const size_t smem = d4c_body_smem_bytes(N);
    const int threads = N / 16;     // one radix-16 butterfly per thread and pass
    p.frame_begin = row0;
    if (n_chunks > 1) {
      for (int c = 0; c < n_chunks; ++c) {
        const int cb = chunk_bound(c), ce = chunk_bound(c + 1);
        cudaStream_t cs = (c & 1) ? alt : stream;
        double *nb = (c & 1) ? d_noise_b : d_noise;
        if ((rc = wb_rng_fill(rng.state, d_skip_mid, d_offsets + ce, (unsigned long long)(ce - cb) * 3ull * N, nb, cs, d_offsets + cb, d_offsets + cb)))
          return rc;
        if (ce > cb) {
          p.frame_begin = cb;
          p.noise = nb;
          p.noise_origin = cb;
          rc = WB_DISPATCH_LOG2(l, 9, 13, {
            if (cudaFuncSetAttribute(d4c_body_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WB_ERR_CUDA;
            WbLaunchScope scope("d4c_body_kernel", cs);
            d4c_body_kernel<L2><<<ce - cb, threads, smem, cs>>>(p);
          });
          if (rc) return rc;
        }
        if (!split) WB_CUDA_CHECK(cudaEventRecord(chunks->ev[c], cs));   // range c may go home
      }
      // the caller's stream continues (randn bookkeeping, later calls) after every chunk
      if (split) {
        WB_CUDA_CHECK(cudaEventRecord(split->join[1], alt));
        WB_CUDA_CHECK(cudaStreamWaitEvent(stream, split->join[1], 0));
      } else {
        for (int c = 1; c < n_chunks; c += 2) WB_CUDA_CHECK(cudaStreamWaitEvent(stream, chunks->ev[c], 0));
      }
    } else {
      rc = WB_DISPATCH_LOG2(l, 9, 13, {
        if (cudaFuncSetAttribute(d4c_body_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WB_ERR_CUDA;
        WB_LAUNCH("d4c_body_kernel", d4c_body_kernel<L2><<<n_rows, threads, smem, stream>>>(p));
      });
      if (rc) return rc;
      if (chunks) for (int c = 0; c < chunks->n; ++c) WB_CUDA_CHECK(cudaEventRecord(chunks->ev[c], stream));
    }
    WB_CUDA_CHECK(cudaGetLastError());
  }
  return rng.advance ? wb_rng_advance(rng.state, d_skip_end, nullptr, stream) : WB_OK;
}
