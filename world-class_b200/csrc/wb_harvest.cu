// Harvest F0 estimation, part 1: everything up to the refined, pruned candidate table.
// (part 2, the contour fixing / smoothing tail, is in wb_harvest_tail.cu)
//
// Reference: /root/reference/src/harvest.cpp
//   ctor :69-103, getSamples :173-181, compute :183-208, getWaveformAndSpectrum :213-248,
//   decimate/FilterForDecimate /root/reference/src/world_matlabfunctions.cpp:27-125,184-210,
//   getFilteredSignal :1261-1305, zeroCrossingEngine :1179-1219,
//   getFourZeroCrossingIntervals :1228-1255, getF0CandidateContour :1098-1143,
//   detectOfficialF0Candidates :1005-1083, overlapF0Candidates :987-1000,
//   refineF0Candidates :932-982, getMeanF0 :883-927 (getBaseIndex :750-757, getMainWindow
//   :762-788, getDiffWindow :794-803, getSpectra :809-842), fixF0 :844-878,
//   removeUnreliableCandidates :708-744, generalBody :1380-1453.
//
// B200 design notes
//  * decimation: the zero-phase 3rd-order IIR is evaluated in parallel chunks, each chunk
//    warmed up over the preceding 400 samples from a zero state (the slowest pole has
//    radius 0.889 -> the state error is < 1e-18 relative after 353 samples).
//  * band-pass bank: the reference convolves by one whole-signal FFT per channel
//    (131072 points for 10 s).  Here the signal is cut into overlap-save blocks of NB
//    samples whose spectra are computed ONCE and shared by all channels; one CTA per
//    channel multiplies by the channel's filter spectrum, inverse-transforms in shared
//    memory and extracts the four kinds of zero crossings straight from shared memory, so
//    the filtered signals (nch x fft_size doubles in the reference) never exist in HBM.
//  * refinement: only <= 6 harmonic bins of each candidate's two spectra are used
//    (fixF0), so a warp evaluates those bins directly instead of two FFTs per candidate.
#include "wb_harvest.h"
#include "wb_fft.cuh"

#include <math.h>
#include <string.h>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// H1: decimation
// ---------------------------------------------------------------------------------------------
struct DecimCoef { double a[3]; double b[2]; };

// world_matlabfunctions.cpp:29-111
bool decimate_coefficients(int r, DecimCoef *c) {
  switch (r) {
    case 11: *c = {{2.450743295230728, -2.06794904601978, 0.59574774438332101}, {0.0026822508007163792, 0.0080467524021491377}}; return true;
    case 12: *c = {{2.4981398605924205, -2.1368928194784025, 0.62187513816221485}, {0.0021097275904709001, 0.0063291827714127002}}; return true;
    case 10: *c = {{2.3936475118069387, -1.9873904075111861, 0.5658879979027055}, {0.0034818622251927556, 0.010445586675578267}}; return true;
    case 9: *c = {{2.3236003491759578, -1.8921545617463598, 0.53148928133729068}, {0.0046331164041389372, 0.013899349212416812}}; return true;
    case 8: *c = {{2.2357462340187593, -1.7780899984041358, 0.49152555365968692}, {0.0063522763407111993, 0.019056829022133598}}; return true;
    case 7: *c = {{2.1225239019534703, -1.6395144861046302, 0.44469707800587366}, {0.0090366882681608418, 0.027110064804482525}}; return true;
    case 6: *c = {{1.9715352749512141, -1.4686795689225347, 0.3893908434965701}, {0.013469181309343825, 0.040407543928031475}}; return true;
    case 5: *c = {{1.7610939654280557, -1.2554914843859768, 0.3237186507788215}, {0.021334858522387423, 0.06400457556716227}}; return true;
    case 4: *c = {{1.4499664446880227, -0.98943497080950582, 0.24578252340690215}, {0.036710750339322612, 0.11013225101796784}}; return true;
    case 3: *c = {{0.95039378983237421, -0.67429146741526791, 0.15412211621346475}, {0.071221945171178636, 0.21366583551353591}}; return true;
    case 2: *c = {{0.041156734567757189, -0.42599112459189636, 0.041037215479961225}, {0.16797464681802227, 0.50392394045406674}}; return true;
    default: *c = {{0.0, 0.0, 0.0}, {0.0, 0.0}}; return false;
  }
}

#define DEC_NFACT 9
#define DEC_CHUNK 32
#define DEC_WARM 400

// tmp1 of decimate() (world_matlabfunctions.cpp:189-193) over new_x of
// getWaveformAndSpectrum (harvest.cpp:222-229), evaluated on the fly from x.
__device__ __forceinline__ double dec_new_x(const double *__restrict__ x, int x_length, int lag, int j) {
  const int k = j - lag;
  return x[k < 0 ? 0 : (k >= x_length ? x_length - 1 : k)];
}
// branch-free: the two source samples are selected with predicates so that the staging loops below can
// keep many loads in flight (interior samples read the same address twice)
__device__ __forceinline__ double dec_tmp1(const double *__restrict__ x, int x_length, int lag, int len1, int i) {
  const bool head = i < DEC_NFACT, tail = i >= DEC_NFACT + len1;
  const int ja = head ? 0 : (tail ? len1 - 1 : i - DEC_NFACT);
  const int jb = head ? DEC_NFACT - i : (tail ? len1 - 2 - (i - (DEC_NFACT + len1)) : ja);
  const double a = dec_new_x(x, x_length, lag, ja), b = dec_new_x(x, x_length, lag, jb);
  return (head || tail) ? 2 * a - b : a;
}

// Both IIR passes work on tiles: a block stages DEC_TILE + DEC_WARM input samples in shared memory
// with coalesced loads, every thread then runs the recurrence over its DEC_CHUNK outputs (plus
// warm-up) out of shared memory, and the results leave through shared memory again.  The
// per-thread stride of DEC_CHUNK doubles is padded to DEC_CHUNK + 1 to stay bank-conflict free.
#define DEC_THREADS 128       /* threads that run the recurrence */
#define DEC_BLOCK 512         /* threads that stage the tiles (load latency, not arithmetic, dominates) */
#define DEC_TILE (DEC_THREADS * DEC_CHUNK)
__device__ __forceinline__ int dec_pad(int r) { return r + (r / DEC_CHUNK); }

// Look-ahead form of the warm-up of w[t] = x[t] + a0 w[t-1] + a1 w[t-2] + a2 w[t-3]: four samples
// per step of the dependency chain (only the state matters during warm-up, so the different
// rounding is irrelevant; the DEC_CHUNK output samples use the reference's own operation order).
struct DecLook {
  double h1, h2, h3;         // impulse response h_0 = 1, h_1.., of the recursive part
  double m2[3], m3[3], m4[3];  // first rows of M^2, M^3, M^4 (M = companion matrix)
};
__device__ __forceinline__ DecLook dec_look_init(const DecimCoef &c) {
  DecLook L;
  const double a0 = c.a[0], a1 = c.a[1], a2 = c.a[2];
  // rows r_k = first row of M^k: r_1 = (a0, a1, a2); r_{k+1} = r_k M
  double r[3] = {a0, a1, a2};
  L.h1 = a0;
  double r2[3] = {r[0] * a0 + r[1], r[0] * a1 + r[2], r[0] * a2};
  L.h2 = r2[0];
  double r3[3] = {r2[0] * a0 + r2[1], r2[0] * a1 + r2[2], r2[0] * a2};
  L.h3 = r3[0];
  double r4[3] = {r3[0] * a0 + r3[1], r3[0] * a1 + r3[2], r3[0] * a2};
  for (int k = 0; k < 3; ++k) { L.m2[k] = r2[k]; L.m3[k] = r3[k]; L.m4[k] = r4[k]; }
  return L;
}
__device__ __forceinline__ void dec_look_step(const DecLook &L, double x1, double x2, double x3, double x4,
                                              double &w0, double &w1, double &w2) {
  const double f2 = fma(L.h1, x1, x2);
  const double f3 = fma(L.h2, x1, fma(L.h1, x2, x3));
  const double f4 = fma(L.h3, x1, fma(L.h2, x2, fma(L.h1, x3, x4)));
  const double n2 = fma(L.m2[0], w0, fma(L.m2[1], w1, fma(L.m2[2], w2, f2)));
  const double n1 = fma(L.m3[0], w0, fma(L.m3[1], w1, fma(L.m3[2], w2, f3)));
  const double n0 = fma(L.m4[0], w0, fma(L.m4[1], w1, fma(L.m4[2], w2, f4)));
  w0 = n0; w1 = n1; w2 = n2;
}

// forward pass: out[i] = FilterForDecimate(tmp1)[i], i in [0, len2)
__global__ void __launch_bounds__(DEC_BLOCK) dec_forward_kernel(const double *__restrict__ x, int x_length, int lag,
                                                                  int len1, int len2, DecimCoef c,
                                                                  double *__restrict__ out) {
  extern __shared__ double dec_smem[];
  double *s_in = dec_smem;                                     // dec_pad(DEC_TILE + DEC_WARM) + 1
  double *s_out = dec_smem + dec_pad(DEC_TILE + DEC_WARM) + 1;  // dec_pad(DEC_TILE) + 1
  const int tile_begin = blockIdx.x * DEC_TILE;
  const int in_begin = tile_begin - DEC_WARM;
#pragma unroll 3
  for (int r = threadIdx.x; r < DEC_TILE + DEC_WARM; r += DEC_BLOCK) {
    const int i = in_begin + r;
    const int ic = min(max(i, 0), len2 - 1);
    const double v = dec_tmp1(x, x_length, lag, len1, ic);
    s_in[dec_pad(r)] = (i >= 0 && i < len2) ? v : 0.0;
  }
  __syncthreads();
  const int begin = tile_begin + threadIdx.x * DEC_CHUNK;
  if (threadIdx.x < DEC_THREADS && begin < len2) {
    const int end = min(len2, begin + DEC_CHUNK);
    double w0 = 0.0, w1 = 0.0, w2 = 0.0;
    const DecLook look = dec_look_init(c);
    int i = max(0, begin - DEC_WARM);
    for (; i + 4 <= begin; i += 4)
      dec_look_step(look, s_in[dec_pad(i - in_begin)], s_in[dec_pad(i + 1 - in_begin)], s_in[dec_pad(i + 2 - in_begin)],
                    s_in[dec_pad(i + 3 - in_begin)], w0, w1, w2);
    for (; i < end; ++i) {
      const double xi = s_in[dec_pad(i - in_begin)];
      const double wt = xi + c.a[0] * w0 + c.a[1] * w1 + c.a[2] * w2;
      if (i >= begin) s_out[dec_pad(i - tile_begin)] = c.b[0] * wt + c.b[1] * w0 + c.b[1] * w1 + c.b[0] * w2;
      w2 = w1; w1 = w0; w0 = wt;
    }
  }
  __syncthreads();
#pragma unroll 4
  for (int r = threadIdx.x; r < DEC_TILE; r += DEC_BLOCK) {
    const int i = tile_begin + r;
    if (i < len2) out[i] = s_out[dec_pad(r)];
  }
}

// backward pass over `fwd` (the reference reverses, filters, reverses) fused with the pick of
// every r-th sample (world_matlabfunctions.cpp:201-207) and the lag removal + zero padding of
// harvest.cpp:231-232,242.  y[m], m in [0, y_length).
__global__ void __launch_bounds__(DEC_BLOCK) dec_backward_kernel(const double *__restrict__ fwd, int len1, int len2,
                                                                   int r, int lag, DecimCoef c, int y_length,
                                                                   double *__restrict__ y,
                                                                   unsigned long long *__restrict__ absmax_bits) {
  extern __shared__ double dec_smem[];
  double *s_in = dec_smem;
  // reversed index u = len2 - 1 - i runs forward in filter time
  const int tile_begin = blockIdx.x * DEC_TILE;
  const int in_begin = tile_begin - DEC_WARM;
#pragma unroll 3
  for (int q = threadIdx.x; q < DEC_TILE + DEC_WARM; q += DEC_BLOCK) {
    const int u = in_begin + q;
    const double v = fwd[len2 - 1 - min(max(u, 0), len2 - 1)];
    s_in[dec_pad(q)] = (u >= 0 && u < len2) ? v : 0.0;
  }
  __syncthreads();
  const int begin = tile_begin + threadIdx.x * DEC_CHUNK;
  double amax = 0.0;  // max |y| over the samples this thread writes (feeds the DC "correction" below)
  if (threadIdx.x < DEC_THREADS && begin < len2) {
  const int end = min(len2, begin + DEC_CHUNK);
  const int nout = len1 / r + 1;
  const int nbeg = r - r * nout + len1;
  double w0 = 0.0, w1 = 0.0, w2 = 0.0;
  const DecLook look = dec_look_init(c);
  int u = max(0, begin - DEC_WARM);
  for (; u + 4 <= begin; u += 4)
    dec_look_step(look, s_in[dec_pad(u - in_begin)], s_in[dec_pad(u + 1 - in_begin)], s_in[dec_pad(u + 2 - in_begin)],
                  s_in[dec_pad(u + 3 - in_begin)], w0, w1, w2);
  for (; u < end; ++u) {
    const double xi = s_in[dec_pad(u - in_begin)];
    const double wt = xi + c.a[0] * w0 + c.a[1] * w1 + c.a[2] * w2;
    if (u >= begin) {
      const double v = c.b[0] * wt + c.b[1] * w0 + c.b[1] * w1 + c.b[0] * w2;
      // tmp1_final[j] with j = len2 - 1 - u; picked when j = i + NFACT - 1, i = nbeg + cnt * r, i < len1 + NFACT
      const int j = len2 - 1 - u;
      const int i = j - DEC_NFACT + 1;
      if (i >= nbeg && i < len1 + DEC_NFACT && (i - nbeg) % r == 0) {
        const int cnt = (i - nbeg) / r;
        const int m = cnt - lag / r;
        if (m >= 0 && m < y_length) { y[m] = v; amax = fmax(amax, fabs(v)); }
      }
    }
    w2 = w1; w1 = w0; w0 = wt;
  }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0 && amax > 0.0) atomicMax(absmax_bits, (unsigned long long)__double_as_longlong(amax));
}

__global__ void copy_kernel(const double *__restrict__ x, int n, double *__restrict__ y,
                            unsigned long long *__restrict__ absmax_bits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (i < n) { v = x[i]; y[i] = v; v = fabs(v); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v > 0.0) atomicMax(absmax_bits, (unsigned long long)__double_as_longlong(v));
}

// harvest.cpp:238-241: `accumulate(y_, y_ + y_length_, 0)` has an int accumulator, i.e. the
// running sum is truncated towards zero after every addition.  If every |y| < 1 the result is
// exactly 0 and y is left alone (the normal case, decided from the max |y| gathered by the
// producer); otherwise the truncating recurrence is replayed sequentially.  One CTA.
__global__ void __launch_bounds__(1024) dc_fix_kernel(double *__restrict__ y, int n,
                                                      const unsigned long long *__restrict__ absmax_bits) {
  __shared__ double s_mean;
  const double amax = __longlong_as_double((long long)*absmax_bits);
  if (amax < 1.0) return;
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < n; ++i) acc = (int)(acc + y[i]);
    double mean_y = acc;
    mean_y /= n;
    s_mean = mean_y;
  }
  __syncthreads();
  const double mean = s_mean;
  for (int i = threadIdx.x; i < n; i += blockDim.x) y[i] -= mean;
}

// ---------------------------------------------------------------------------------------------
// H3/H4: overlap-save block spectra and per-channel filter spectra
// ---------------------------------------------------------------------------------------------
template <int LOG2NB>
__global__ void __launch_bounds__(512) yspec_kernel(const double *__restrict__ y, int y_length, int V, int h_max,
                                                    const cplx *__restrict__ tw, cplx *__restrict__ Yb) {
  extern __shared__ double2 smem_raw[];
  cplx *S = smem_raw;
  double *W = reinterpret_cast<double *>(S);
  constexpr int NB = 1 << LOG2NB, NC = NB / 2;
  const int b = blockIdx.x;
  const long long s_b = (long long)b * V + 1 - h_max;
  for (int i = threadIdx.x; i < NB; i += blockDim.x) {
    const long long n = s_b + i;
    W[wb_didx(i)] = (n >= 0 && n < y_length) ? y[n] : 0.0;
  }
  __syncthreads();
  cplx *dst = Yb + (size_t)b * (NC + 1);
  wb_rfft_t<1, LOG2NB - 1>(S, tw, [&](int k, cplx X) { dst[k] = X; });
}

// getFilteredSignal (harvest.cpp:1264-1274): Nuttall(2h+1) x cos(2 pi bf i / fs), then r2c
template <int LOG2NB>
__global__ void __launch_bounds__(512) filter_spec_kernel(const double *__restrict__ boundary_f0, const int *__restrict__ half_len,
                                                          double actual_fs, const cplx *__restrict__ tw,
                                                          cplx *__restrict__ Hc) {
  extern __shared__ double2 smem_raw[];
  cplx *S = smem_raw;
  double *W = reinterpret_cast<double *>(S);
  constexpr int NB = 1 << LOG2NB, NC = NB / 2;
  const int c = blockIdx.x;
  const double bf = boundary_f0[c];
  const int h = half_len[c];
  const int flen = 2 * h + 1;
  for (int i = threadIdx.x; i < NB; i += blockDim.x) {
    double v = 0.0;
    if (i < flen) {
      const double tmp = i / (flen - 1.0);
      const double win = 0.355768 - 0.487396 * cos(2.0 * WB_PI * tmp) + 0.144232 * cos(4.0 * WB_PI * tmp) -
                         0.012604 * cos(6.0 * WB_PI * tmp);
      v = win * cos(2 * WB_PI * bf * (i - h) / actual_fs);
    }
    W[wb_didx(i)] = v;
  }
  __syncthreads();
  cplx *dst = Hc + (size_t)c * (NC + 1);
  wb_rfft_t<1, LOG2NB - 1>(S, tw, [&](int k, cplx X) { dst[k] = X; });
}

// ---------------------------------------------------------------------------------------------
// H5: per-channel band-pass + zero crossings
// ---------------------------------------------------------------------------------------------
struct ChanParams {
  const cplx *Yb; const cplx *Hc; const int *half_len;
  int n_blocks; int NB; int log2nc; int V; int h_max; int y_length;
  const cplx *tw;
  double *seg_edges;  // [nch][4][n_blocks][bcap]
  int *seg_count;     // [nch][4][n_blocks]
  int bcap;
};

#define CH_THREADS 256
#define CH_MASK_WORDS 5   /* flags of up to 80 samples per thread (the run is V / 256 | 1 <= 65 samples for blocks of 2^14) */

// flags of sample n for the four zero-crossing kinds (harvest.cpp:1179-1255):
// bit 0: negative-going of f, 1: of -f, 2: of d = f[n+1]-f[n] (peaks), 3: of -d (dips)
__device__ __forceinline__ unsigned ch_flags(double a, double b, double c, int n, int y_length) {
  unsigned m = 0;
  if (n < y_length - 1) {
    if (0.0 < a && b <= 0.0) m |= 1u;
    if (a < 0.0 && b >= 0.0) m |= 2u;
  }
  if (n < y_length - 2) {
    const double d0 = b - a, d1 = c - b;
    if (0.0 < d0 && d1 <= 0.0) m |= 4u;
    if (d0 < 0.0 && d1 >= 0.0) m |= 8u;
  }
  return m;
}

// One CTA per (overlap-save block, channel): band-pass by the convolution theorem
// (harvest.cpp:1277-1299) and order-preserving extraction of the fine zero-crossing edges of the
// block's V output samples straight from shared memory.
template <int LOG2NB>
__global__ void __launch_bounds__(CH_THREADS, 3) channel_kernel(ChanParams p) {
  extern __shared__ double2 smem_raw[];
  __shared__ int s_wsum[4][CH_THREADS / 32];
  cplx *S = smem_raw;
  double *W = reinterpret_cast<double *>(S);
  constexpr int NC = (1 << LOG2NB) / 2;
  const int b = blockIdx.x, c = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = p.half_len[c];
  const cplx *H = p.Hc + (size_t)c * (NC + 1);
  const cplx *Y = p.Yb + (size_t)b * (NC + 1);
  // (one radix-16 butterfly per thread and pass when the block spectrum has 4096 complex points: warp-local late passes)
  wb_irfft_t<-1, LOG2NB - 1, 16, (NC / 16 <= CH_THREADS)>(S, p.tw, [&](int k) {
    const cplx yv = Y[k], hv = H[k];
    return make_double2(yv.x * hv.x - yv.y * hv.y, yv.x * hv.y + yv.y * hv.x);
  });
  // filtered[n] = seg[n + 1 + h - s_b], s_b = b V + 1 - h_max   (delay compensation :1297-1299)
  const int n0 = b * p.V;
  const int off = h + p.h_max - n0;
  const int n_end = min(n0 + p.V, p.y_length);
  // each thread owns a contiguous run of samples (odd length: conflict-free shared-memory reads)
  int chunk = (p.V + CH_THREADS - 1) / CH_THREADS;
  chunk |= 1;   // (<= 16 * CH_MASK_WORDS: V <= 2^14)
  const int my_begin = n0 + tid * chunk, my_end = min(n_end, my_begin + chunk);
  // (a three-sample window slides over the run: one shared-memory read and one padded-index computation per
  // sample instead of three; the flags of the run are kept as a bit mask, so the second pass only visits the
  // samples that have a crossing -- a few per thread)
  int cnt[4] = {0, 0, 0, 0};
  unsigned long long mask[CH_MASK_WORDS];
#pragma unroll
  for (int w = 0; w < CH_MASK_WORDS; ++w) mask[w] = 0ull;
  if (my_begin < my_end) {
    double a = W[wb_didx(my_begin + off)], bb = W[wb_didx(my_begin + 1 + off)];
#pragma unroll
    for (int w = 0; w < CH_MASK_WORDS; ++w) {
      if (my_begin + w * 16 >= my_end) break;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int n = my_begin + w * 16 + q;
        if (n < my_end) {
          const double cc = W[wb_didx(n + 2 + off)];
          const unsigned m = ch_flags(a, bb, cc, n, p.y_length);
          mask[w] |= (unsigned long long)m << (4 * q);
          cnt[0] += m & 1u; cnt[1] += (m >> 1) & 1u; cnt[2] += (m >> 2) & 1u; cnt[3] += (m >> 3) & 1u;
          a = bb; bb = cc;
        }
      }
    }
  }
  int pre[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    int incl = cnt[t];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_wsum[t][warp] = incl;
    pre[t] = incl - cnt[t];
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    int total = 0;
    for (int w = 0; w < CH_THREADS / 32; ++w) {
      const int v = s_wsum[t][w];
      if (w < warp) pre[t] += v;
      total += v;
    }
    if (tid == 0) p.seg_count[((size_t)c * 4 + t) * p.n_blocks + b] = min(total, p.bcap);
  }
  double *dst = p.seg_edges + ((size_t)c * 4 * p.n_blocks + b) * p.bcap;
  const size_t tstride = (size_t)p.n_blocks * p.bcap;
#pragma unroll
  for (int w = 0; w < CH_MASK_WORDS; ++w) {
    unsigned long long left = mask[w];
    while (left) {
      const int q = (__ffsll((long long)left) - 1) >> 2;
      const unsigned m = (unsigned)(left >> (4 * q)) & 15u;
      left &= ~(15ull << (4 * q));
      const int n = my_begin + w * 16 + q;
      const double a = W[wb_didx(n + off)], bb = W[wb_didx(n + 1 + off)];
      if (m & 3u) {
        const double fine = (n + 1) - a / (bb - a);
        if (m & 1u) { if (pre[0] < p.bcap) dst[pre[0]] = fine; ++pre[0]; }
        if (m & 2u) { if (pre[1] < p.bcap) dst[tstride + pre[1]] = fine; ++pre[1]; }
      }
      if (m & 12u) {
        const double cc = W[wb_didx(n + 2 + off)];
        const double d0 = bb - a, d1 = cc - bb;
        const double fine = (n + 1) - d0 / (d1 - d0);
        if (m & 4u) { if (pre[2] < p.bcap) dst[2 * tstride + pre[2]] = fine; ++pre[2]; }
        if (m & 8u) { if (pre[3] < p.bcap) dst[3 * tstride + pre[3]] = fine; ++pre[3]; }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// zeroCrossingEngine's interval sequences (harvest.cpp:1208-1211) and getF0CandidateContour's four
// interp1 calls (harvest.cpp:1098-1143, world_matlabfunctions.cpp:136-182), evaluated in the INVERSE
// direction: instead of every frame searching the interval list (a chain of dependent loads), every
// interval writes the frames it covers.  One CTA per (channel, kind).
// ---------------------------------------------------------------------------------------------
struct IntervalParams {
  const double *seg_edges; const int *seg_count; int n_blocks; int bcap;
  int *ecount; int ecap; double fs;
  int f0_length; int frame_period;
  const double *t_tab;  // [f0_length] frame times
  double *contour;  // interpolated interval frequency per (channel, kind, frame), see iv_contour_index
};

// Layout of the four interpolated contours: tile-major, [frame / 32][channel][kind][frame % 32], so that the
// candidate kernel (one CTA per 32 frames) streams one contiguous block.
__device__ __forceinline__ size_t iv_contour_index(int ct, int n_ct, int frame) {
  return ((size_t)(frame >> 5) * n_ct + ct) * 32 + (frame & 31);
}

// smallest frame i in [0, L] with t_i >= x; t_tab[i] = i * frame_period / 1000.0 (the reference's own
// expression, tabulated so that the frame loop has no division for it)
__device__ __forceinline__ int iv_first_frame(double x, double frames_per_second, const double *__restrict__ t_tab, int L) {
  int g = static_cast<int>(ceil(x * frames_per_second));
  g = wb_max_i(0, wb_min_i(L, g));
  while (g > 0 && t_tab[g - 1] >= x) --g;
  while (g < L && t_tab[g] < x) ++g;
  return g;
}

__global__ void frame_time_kernel(int n, int frame_period, double *__restrict__ t_tab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t_tab[i] = i * frame_period / 1000.0;
}

#define IV_THREADS 512
#define IV_MAX_BLOCKS 1024   /* overlap-save blocks per utterance: ~15 min of audio at the 8 kHz analysis rate */
#define IV_PER_ITER ((IV_THREADS / 32) * 31)
__global__ void __launch_bounds__(IV_THREADS) interval_kernel(IntervalParams p) {
  __shared__ int s_off[IV_MAX_BLOCKS + 1];
  // channel * 4 + kind; the high channels have the most zero crossings: schedule them first
  const int ct = gridDim.x - 1 - blockIdx.x;
  const int *cnt = p.seg_count + (size_t)ct * p.n_blocks;
  const int nb = p.n_blocks;
  if (threadIdx.x < 32) {
    // exclusive offsets of the per-block edge runs (n_blocks <= IV_MAX_BLOCKS): the ordered edge list of
    // this (channel, kind) is their concatenation.  Lane l owns the blocks [l per, (l + 1) per).
    const int lane = threadIdx.x;
    const int per = (nb + 31) >> 5, b0 = lane * per, b1 = min(b0 + per, nb);
    int sum = 0;
    for (int b = b0; b < b1; ++b) sum += cnt[b];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    int run = incl - sum;
    for (int b = b0; b < b1; ++b) { s_off[b] = run; run += cnt[b]; }
    if (lane == 31) { s_off[nb] = incl; p.ecount[ct] = min(incl, p.ecap); }
  }
  __syncthreads();
  const int total = min(s_off[nb], p.ecap);
  const int ni = total < 2 ? 0 : total - 1;  // number of intervals
  const double *seg = p.seg_edges + (size_t)ct * p.n_blocks * p.bcap;
  // edge k of the concatenated list: last block b with s_off[b] <= k
  auto edge = [&](int k) -> double {
    int lo = 0, hi = nb - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_off[mid] <= k) lo = mid; else hi = mid - 1;
    }
    return seg[(size_t)lo * p.bcap + (k - s_off[lo])];
  };
  double *out = p.contour;
  const int n_ct = gridDim.x;
  const double fs = p.fs;
  const int L = p.f0_length;
  const double *t_tab = p.t_tab;
  const double frames_per_second = 1000.0 / p.frame_period;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // Lanes 1..31 of a warp own 31 consecutive intervals; lane 0 recomputes the interval before them, so
  // every owner gets its left knot (and the first frame at or after it) with one shuffle.
  for (int k0 = 0; k0 < ni; k0 += IV_PER_ITER) {
    const int k = k0 + warp * 31 + lane - 1;
    double x1 = 0.0, y1 = 0.0;
    int f1 = L;
    if (k >= 0 && k < ni) {
      // zeroCrossingEngine (harvest.cpp:1208-1211): interval value and location
      const double e0 = edge(k), e1 = edge(k + 1);
      y1 = fs / (e1 - e0);
      x1 = (e0 + e1) / 2.0 / fs;
      if (k < ni - 1) f1 = iv_first_frame(x1, frames_per_second, t_tab, L);
    }
    const double x0 = __shfl_up_sync(0xffffffffu, x1, 1);
    const double y0 = __shfl_up_sync(0xffffffffu, y1, 1);
    const int f0 = __shfl_up_sync(0xffffffffu, f1, 1);
    // interp1's histc puts frame t between knots k - 1 and k when loc[k-1] <= t < loc[k]; the index is
    // clamped to [1, ni - 1], i.e. the first and last knot pairs extrapolate (needs ni > 2, checked by
    // the consumer exactly like harvest.cpp:1113-1117)
    if (lane > 0 && k >= 1 && k < ni && ni > 2) {
      const int i_lo = (k > 1) ? f0 : 0;
      // interp1 divides by the knot distance for every frame (world_matlabfunctions.cpp:168-176); it is a loop
      // invariant here, so its reciprocal is formed once (the quotient can differ in the last bit)
      const double inv_dx = 1.0 / (x1 - x0), dy = y1 - y0;
      for (int i = i_lo; i < f1; ++i) {
        const double s = (t_tab[i] - x0) * inv_dx;
        out[iv_contour_index(ct, n_ct, i)] = y0 + s * dy;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// getF0CandidateContour's gating (harvest.cpp:1126-1142) + detectOfficialF0Candidates
// (harvest.cpp:1005-1083).  One CTA per 32 frames: the warps sweep the channels (coalesced reads of
// the four contours), the gated raw candidates of the tile go to shared memory, then one thread
// per frame runs the reference's run-length scan over the channels and appends the frame's
// candidates to the refinement work list.
// ---------------------------------------------------------------------------------------------
struct CandParams {
  const double *contour; const int *ecount; const double *boundary_f0;
  int nch; int f0_length; double f0_floor; double f0_ceil;
  double *raw;       // [nch][f0_length]
  double *own;       // [f0_length][own_cap]
  int own_cap;
  int *nc_max;       // [0] = max candidates per frame, [1] = work count
  int *work;         // frame * 32 + j of every own candidate
};

#define CD_FRAMES 32
#define CD_PITCH 33
#define CD_THREADS 512
__global__ void __launch_bounds__(CD_THREADS) candidate_kernel(CandParams p) {
  extern __shared__ double cd_tile[];  // [nch][CD_PITCH]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int L = p.f0_length;
  const int f = blockIdx.x * CD_FRAMES + lane;
#pragma unroll 4
  for (int c = warp; c < p.nch; c += CD_THREADS / 32) {
    const int *cnt = p.ecount + c * 4;
    bool ok = true;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int ni = cnt[t] < 2 ? 0 : cnt[t] - 1;
      if (ni - 2 <= 0) ok = false;
    }
    double v = 0.0;
    if (ok && f < L) {
      const double *src = p.contour + iv_contour_index(c * 4, p.nch * 4, f);
      const double v0 = src[0], v1 = src[32], v2 = src[64], v3 = src[96];
      const double bf = p.boundary_f0[c];
      const double upper = bf * 1.1, lower = bf * 0.9;
      v = (v0 + v1 + v2 + v3) / 4.0;
      if (v > upper || v < lower || v > p.f0_ceil || v < p.f0_floor) v = 0.0;
    }
    if (f < L) p.raw[(size_t)c * L + f] = v;
    cd_tile[c * CD_PITCH + lane] = v;
  }
  __syncthreads();
  if (warp != 0 || f >= L) return;
  double *dst = p.own + (size_t)f * p.own_cap;
  const int nch = p.nch;
  int count = 0;
  int prev = 0, st = 0;
  double acc = 0.0;
  for (int j = 1; j < nch; ++j) {
    // vuv[0] = vuv[nch-1] = 0
    const double r = cd_tile[j * CD_PITCH + lane];
    const int cur = (j == nch - 1) ? 0 : (r > 0 ? 1 : 0);
    if (cur - prev == 1) { st = j; acc = 0.0; }
    if (cur == 1) acc += r;
    if (cur - prev == -1) {
      const int ed = j;
      if (ed - st >= 10 && count < p.own_cap) dst[count++] = acc / (ed - st);
    }
    prev = cur;
  }
  for (int k = count; k < p.own_cap; ++k) dst[k] = 0.0;
  if (count > 0) {
    atomicMax(p.nc_max, count);
    const int at = atomicAdd(p.nc_max + 1, count);
    for (int j = 0; j < count; ++j) p.work[at + j] = f * 32 + j;
  }
}

// ---------------------------------------------------------------------------------------------
// refineF0Candidates / getMeanF0 / fixF0 (harvest.cpp:844-982): one warp per candidate
// ---------------------------------------------------------------------------------------------
struct RefineParams {
  const double *y; int y_length; double actual_fs;
  double f0_floor, f0_ceil; int frame_period;
  const int *work; const int *nc_and_count;   // own candidates (frame * 32 + j); [0] = nc, [1] = their number
  const double *own; int own_cap; int f0_length;
  int max_candidates;
  double *cand; double *score;                // zero-initialised [f0_length][max_candidates] tables
  const cplx *tw[16];   // twiddle tables by log2(fft_size)
  int max_wlen;         // shared window buffer length per warp
};

#define RF_WARPS 8

// Blackman main window at angle cosine c (harvest.cpp:770-774), cos(2t) = 2 cos(t)^2 - 1
__device__ __forceinline__ double rf_window(double c) { return 0.42 + 0.5 * c + 0.08 * (2.0 * c * c - 1.0); }

// ---------------------------------------------------------------------------------------------
// The refinement is a small dense contraction and runs on the fp64 tensor pipe.  One warp per OWN candidate: its
// seven copies (frames src-3 .. src+3, harvest.cpp:987-1000) have the same f0, hence the same window length, the
// same FFT size and the same harmonic bins -- only the waveform slice differs.  With
//   A[(copy, window)][i] = window_copy(i) * y[first_copy + i]      (7 x 2 rows: main and differentiated window)
//   B[i][(h, cos | sin)] = e^{+2 pi i idx_h i / fft_size}           (6 x 2 columns, shared by the seven copies)
// the two spectra of every copy at the <= 6 bins fixF0 uses (harvest.cpp:809-878) are C = A B, accumulated by
// mma.sync.m8n8k4.f64: M tile 0 = main windows (row = copy), M tile 1 = differentiated windows, N tile 0 = cosines
// (column = harmonic), N tile 1 = sines, K = four window samples per step.  Lane (g, t) = (lane / 4, lane % 4)
// contributes A[g][i0 + t] and B[i0 + t][g]: ONE waveform load, one window evaluation (a rotation recurrence:
// the window angle advances by exactly 2 pi / len per sample; end rules of harvest.cpp:794-803) and one twiddle gather per lane and step feed four MMAs = 1024
// multiply-adds, where scalar code issues 24 DFMA per lane for 24 x 32.  The accumulator fragment leaves lane
// (g, t) with main and differentiated spectra of copy g at harmonics 2t, 2t + 1: fixF0's per-harmonic terms need
// no exchange, the final sums over the harmonics run in the reference's order.
// Work items are handed out through a global counter (windows are 31 .. 600 samples long).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void wb_dmma_8x8x4(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Per-candidate constants of the refinement, computed by ONE THREAD per own candidate (refine_setup_kernel) instead of
// redundantly by the 32 lanes of the warp that accumulates the candidate: window half length, FFT size, harmonic bins,
// the first window sample of each of the seven copies, the window's one-sample rotation and -- on grids where the
// windows are symmetric about a sample -- the starting rotations of the four pair sequences and the window angles of
// the four samples with an end rule.
struct RefineSetup {
  double f0, s1, c1;            // candidate; sin / cos of 2 pi / window length
  double init_sn[4], init_cs[4];  // symmetric case: sin / cos of 2 pi t / window length, t = 0 .. 3
  double edge_sn[4], edge_cs[4];  // symmetric case: window angle of samples 0, 1, 2, 2 hw (the reference's expression)
  int src, own_j, hw, log2fft, nh, symmetric, active, pad;
  int idx[8];                   // harmonic bins (6 used)
  int basic_index[8];           // first window sample of copy o (7 used)
};

// eight threads per candidate: thread k takes copy k and harmonic k, threads 0 .. 3 one starting rotation and one
// edge-sample angle each
__global__ void __launch_bounds__(128) refine_setup_kernel(RefineParams p, RefineSetup *__restrict__ out) {
  const int lane = threadIdx.x & 31, k = lane & 7;
  const int n_own = p.nc_and_count[1];
  const double fs = p.actual_fs;
  const double two_pi = 2.0 * WB_PI;
  // a warp takes four candidates per step (the loop bound is warp-uniform: the ballots below need the whole warp)
  const int warp_first = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4;
  const int warp_stride = ((gridDim.x * blockDim.x) >> 5) * 4;
  for (int c0 = warp_first; c0 < n_own; c0 += warp_stride) {
  const int c = c0 + (lane >> 3);
  const bool live = c < n_own;
  const int item = live ? p.work[c] : 0;
  const int src = item >> 5, own_j = item & 31;
  const double current_f0 = live ? p.own[(size_t)src * p.own_cap + own_j] : 100.0;
  const int hw = static_cast<int>(1.5 * fs / current_f0 + 1.0);
  const double window_length_in_time = (2.0 * hw + 1.0) / fs;
  const int log2fft = 2 + (31 - __clz(2 * hw + 1));  // 2 + int(log2(len)), len odd
  const int fft_size = 1 << log2fft;
  const int i_c = hw + 1;
  // copy k (row 7 is padding)
  const int frame = (k <= 3) ? src + k : src - (k - 3);
  const bool act = live && k < 7 && frame >= 0 && frame < p.f0_length;
  const double current_position = frame * p.frame_period / 1000.0;
  const double base_time0 = (-hw + 0) / fs;
  const int bi = wb_round((current_position + base_time0) * fs + 0.001);
  // the reference's own expression for the window argument of the centre sample: zero when it is on the grid
  const double centre = ((bi + i_c) - 1.0) / fs - current_position;
  const unsigned grp = 0xffu << (lane & 24);
  const unsigned act_bits = (__ballot_sync(0xffffffffu, act) & grp) >> (lane & 24);
  const unsigned off_bits = (__ballot_sync(0xffffffffu, act && !(fabs(centre) * fs < 1e-6)) & grp) >> (lane & 24);
  const int symmetric = off_bits == 0u ? 1 : 0;
  // the first active copy supplies the edge-sample angles of the symmetric case (they are the same for every copy there)
  const int ref_lane = (lane & 24) + (act_bits ? __ffs(act_bits) - 1 : 0);
  const int ref_bi = __shfl_sync(0xffffffffu, bi, ref_lane);
  const double ref_pos = __shfl_sync(0xffffffffu, current_position, ref_lane);
  if (!live) continue;
  RefineSetup *r = out + c;
  r->idx[k] = k < 6 ? wb_round(current_f0 * fft_size / fs * (k + 1)) : 0;
  r->basic_index[k] = bi;
  if (k < 4) {
    double sn = 0.0, cs = 1.0, esn = 0.0, ecs = 1.0;
    if (symmetric) {
      sincos(two_pi * k / (2 * hw + 1), &sn, &cs);
      const int i = k < 3 ? k : 2 * hw;
      const double tmp = ((ref_bi + i) - 1.0) / fs - ref_pos;
      sincos(two_pi * tmp / window_length_in_time, &esn, &ecs);
    }
    r->init_sn[k] = sn; r->init_cs[k] = cs; r->edge_sn[k] = esn; r->edge_cs[k] = ecs;
  } else if (k == 4) {
    double s1, c1;
    sincos(two_pi / (2 * hw + 1), &s1, &c1);
    r->f0 = current_f0; r->s1 = s1; r->c1 = c1;
  } else if (k == 5) {
    r->src = src; r->own_j = own_j; r->hw = hw; r->log2fft = log2fft;
    r->nh = wb_min_i(static_cast<int>(fs / 2.0 / current_f0), 6);
    r->symmetric = symmetric; r->active = (int)act_bits; r->pad = 0;
  }
  }
}

__global__ void __launch_bounds__(RF_WARPS * 32, 3) refine_mma_kernel(RefineParams p, const RefineSetup *__restrict__ setup,
                                                                       int *__restrict__ ticket) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int nc = p.nc_and_count[0];
  const int n_own = p.nc_and_count[1];
  const double fs = p.actual_fs;
  const double two_pi = 2.0 * WB_PI;
  // The ticket of the NEXT candidate is requested at the top of an iteration and used at its end (the atomic is a
  // round trip to L2).
  int c = 0;
  if (lane == 0) c = atomicAdd(ticket, 1);
  c = __shfl_sync(0xffffffffu, c, 0);
  while (c < n_own) {
    int tick = 0;
    if (lane == 0) tick = atomicAdd(ticket, 1);
    const RefineSetup *rs = setup + c;
    const int src = rs->src, own_j = rs->own_j;
    const double current_f0 = rs->f0;
    const int hw = rs->hw;
    const int len = 2 * hw + 1;
    const double window_length_in_time = (2.0 * hw + 1.0) / fs;
    const int log2fft = rs->log2fft;
    const int fft_size = 1 << log2fft;
    const int nh = rs->nh;
    const cplx *T = p.tw[log2fft];
    const int mask = fft_size - 1;
    // row of this lane: copy o = g (row 7 is padding)
    const int o = g;
    const int frame = (o <= 3) ? src + o : src - (o - 3);
    const bool row_active = (rs->active >> o) & 1;
    const double current_position = frame * p.frame_period / 1000.0;
    const int basic_index = rs->basic_index[o];
    // column of this lane: harmonic g (columns 6, 7 are padding)
    const int idx_col = rs->idx[g];
    const double s1 = rs->s1, c1 = rs->c1;
    // rotation by four samples from the one-sample rotation (two angle doublings; 1 - 2 sin^2 keeps the cosine accurate)
    const double s2 = 2.0 * s1 * c1, c2a = 1.0 - 2.0 * s1 * s1;
    const double sg = 2.0 * s2 * c2a, cg = 1.0 - 2.0 * s2 * s2;
    const double dw_a = 0.5 * s1, dw_b = 0.16 * (2.0 * s1 * c1);   // interior samples of the differentiated window (harvest.cpp:794-803) in closed form: with w(a) = 0.42 + 0.5 cos a + 0.08 cos 2a, -(w(a + d) - w(a - d)) / 2 = sin a (0.5 sin d + 0.16 sin 2d cos a)
    double mc0 = 0.0, mc1 = 0.0, ms0 = 0.0, ms1 = 0.0, dc0 = 0.0, dc1 = 0.0, ds0 = 0.0, ds1 = 0.0;
    const int y_first = basic_index - 1, y_last = p.y_length - 1;
    const int i_c = hw + 1;   // the sample the window is centred on
    const bool symmetric = rs->symmetric != 0;
    // main and differentiated window times y for window sample i at window angle (sn, cs) (harvest.cpp:762-803)
    auto sample_at = [&](int i, double sn, double cs, double &vm, double &vd) {
      const double yv = row_active ? p.y[wb_max_i(0, wb_min_i(y_last, y_first + i))] : 0.0;
      double dwin = sn * fma(dw_b, cs, dw_a);
      if (i == 0) dwin = -rf_window(cs * c1 - sn * s1) / 2.0;            // -w[1] / 2
      else if (i == len - 1) dwin = rf_window(cs * c1 + sn * s1) / 2.0;  // w[len - 2] / 2
      vm = rf_window(cs) * yv;
      vd = dwin * yv;
    };
    if (symmetric) {
      // Frame times on whole decimated samples (actual_fs a multiple of 1000 Hz): the window is even and its
      // differentiated version odd about the centre sample, so the samples at +j and -j share one twiddle
      // and K halves: with ys = y_j + y_-j, yd = y_j - y_-j,
      //   main += w_j (ys cos + i yd sin),   diff += d_j (yd cos + i ys sin),   spectra referred to the centre sample
      // (fixF0 only uses |main|^2 and Im(conj(main) diff), which a common phase does not change).
      // Pairs j = 1 .. hw - 2 (samples 3 .. 2 hw - 1); the five samples without a partner or with an end rule
      // (0, 1, 2, hw + 1, 2 hw) follow in two more steps.
      // (pair index j = 0 is the centre sample itself: window 1, differentiated window 0, no partner)
      double sn = rs->init_sn[t], cs = rs->init_cs[t];
      const int yc = y_first + i_c;
      const int n_pairs = hw - 2;
      int j = t;
      double yp_next = 0.0, ym_next = 0.0;
      if (row_active && j <= n_pairs) {
        yp_next = p.y[wb_max_i(0, wb_min_i(y_last, yc + j))];
        ym_next = j > 0 ? p.y[wb_max_i(0, wb_min_i(y_last, yc - j))] : 0.0;
      }
      cplx w_next = __ldg(&T[(idx_col * j) & mask]);
      for (int j0 = 0; j0 <= n_pairs; j0 += 4, j += 4) {
        const double yp = yp_next, ym = ym_next;
        const cplx w = w_next;
        const int jn = j + 4;
        yp_next = 0.0; ym_next = 0.0;
        if (row_active && jn <= n_pairs) {
          yp_next = p.y[wb_max_i(0, wb_min_i(y_last, yc + jn))];
          ym_next = p.y[wb_max_i(0, wb_min_i(y_last, yc - jn))];
        }
        w_next = __ldg(&T[(idx_col * jn) & mask]);
        const double w_here = fma(cs, fma(0.16, cs, 0.5), 0.34);
        const double dwin = sn * fma(dw_b, cs, dw_a);
        const double ys = yp + ym, yd = yp - ym;     // (both zero past the last pair)
        wb_dmma_8x8x4(mc0, mc1, w_here * ys, w.x);
        wb_dmma_8x8x4(ms0, ms1, w_here * yd, w.y);
        wb_dmma_8x8x4(dc0, dc1, dwin * yd, w.x);
        wb_dmma_8x8x4(ds0, ds1, dwin * ys, w.y);
        const double c2 = fma(cs, cg, -(sn * sg));
        sn = fma(sn, cg, cs * sg);
        cs = c2;
      }
      {
        // the four samples with an end rule or without a partner: i = 0, 1, 2 and 2 hw, one per lane of the row's group
        // (hw >= 2 always: f0 <= f0_ceil * 1.1 well below 1.5 fs / 2; for tiny windows some of these coincide)
        const int i = t < 3 ? t : 2 * hw;
        bool use = i < len && i != i_c;
        if (t == 3) use = use && i > 2;
        double vm = 0.0, vd = 0.0;
        if (use) sample_at(i, rs->edge_sn[t], rs->edge_cs[t], vm, vd);
        const cplx w = __ldg(&T[(idx_col * (i - i_c)) & mask]);
        wb_dmma_8x8x4(mc0, mc1, vm, w.x);
        wb_dmma_8x8x4(ms0, ms1, vm, w.y);
        wb_dmma_8x8x4(dc0, dc1, vd, w.x);
        wb_dmma_8x8x4(ds0, ds1, vd, w.y);
      }
    } else {
      double sn, cs;
      {
        const double tmp = ((basic_index + t) - 1.0) / fs - current_position;
        sincos(two_pi * tmp / window_length_in_time, &sn, &cs);
      }
      double y_next = (row_active && t < len) ? p.y[wb_max_i(0, wb_min_i(y_last, y_first + t))] : 0.0;
      cplx w_next = __ldg(&T[(idx_col * t) & mask]);
      for (int i = t; i < len + t; i += 4) {   // (every lane makes the same number of steps: the MMAs are warp-wide)
        const double yv = y_next;
        const cplx w = w_next;
        const int in = i + 4;
        y_next = (row_active && in < len) ? p.y[wb_max_i(0, wb_min_i(y_last, y_first + in))] : 0.0;
        w_next = __ldg(&T[(idx_col * in) & mask]);
        double a_main = 0.0, a_diff = 0.0;
        if (i < len) {
          const double w_here = fma(cs, fma(0.16, cs, 0.5), 0.34);  // = rf_window(cs)
          double dwin = sn * fma(dw_b, cs, dw_a);
          if (i == 0) dwin = -rf_window(cs * c1 - sn * s1) / 2.0;            // -w[1] / 2
          else if (i == len - 1) dwin = rf_window(cs * c1 + sn * s1) / 2.0;  // w[len - 2] / 2
          a_main = w_here * yv;
          a_diff = dwin * yv;
        }
        wb_dmma_8x8x4(mc0, mc1, a_main, w.x);
        wb_dmma_8x8x4(ms0, ms1, a_main, w.y);
        wb_dmma_8x8x4(dc0, dc1, a_diff, w.x);
        wb_dmma_8x8x4(ds0, ds1, a_diff, w.y);
        const double c2 = fma(cs, cg, -(sn * sg));
        sn = fma(sn, cg, cs * sg);
        cs = c2;
      }
    }
    // fixF0 (harvest.cpp:844-878) for harmonics 2t and 2t + 1 of copy g; spectra are conjugated by the reference
    // (harvest.cpp:829-841): main = (mr, -mi), diff = (dr, -di)
    double inst[2], amp[2], dev[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int hh = 2 * t + e;
      const double m_re = e ? mc1 : mc0, m_im = -(e ? ms1 : ms0), d_re = e ? dc1 : dc0, d_im = -(e ? ds1 : ds0);
      const int my_idx = wb_round(current_f0 * fft_size / fs * (hh + 1));
      const double power = m_re * m_re + m_im * m_im;
      const double num_i = m_re * d_im - m_im * d_re;
      inst[e] = (power == 0.0) ? 0.0 : static_cast<double>(my_idx) * fs / fft_size + num_i / power * fs / 2.0 / WB_PI;
      amp[e] = sqrt(power);
      dev[e] = fabs((inst[e] / (hh + 1.0) - current_f0) / current_f0);
    }
    // the sums over the harmonics in the reference's order (hh = 0 .. nh - 1): lane t adds its two terms to the running
    // sums and hands them to lane t + 1 of the row's group; lane 2 ends up with the totals
    double numerator = 0.0, denominator = 0.0, score = 0.0;
#pragma unroll
    for (int hop = 0; hop < 3; ++hop) {
      if (hop > 0) {
        numerator = __shfl_up_sync(0xffffffffu, numerator, 1, 4);
        denominator = __shfl_up_sync(0xffffffffu, denominator, 1, 4);
        score = __shfl_up_sync(0xffffffffu, score, 1, 4);
      }
      if (t == hop) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int hh = 2 * hop + e;
          if (hh < nh) {
            numerator += amp[e] * inst[e];
            denominator += amp[e] * (hh + 1.0);
            score += dev[e];
          }
        }
      }
    }
    if (t == 2 && row_active) {
      double refined = numerator / (denominator + WB_SAFEGUARD);
      double sc = 1.0 / (score / nh + WB_SAFEGUARD);
      if (refined < p.f0_floor || refined > p.f0_ceil || sc < 2.5) { refined = 0.0; sc = 0.0; }
      const size_t at = (size_t)frame * p.max_candidates + (o * nc + own_j);
      p.cand[at] = refined;
      p.score[at] = sc;
    }
    c = __shfl_sync(0xffffffffu, tick, 0);
  }
}

// ---------------------------------------------------------------------------------------------
// removeUnreliableCandidates (harvest.cpp:708-744): one thread per (frame, slot)
// ---------------------------------------------------------------------------------------------
// RM_FRAMES frames per CTA: the candidate rows of the tile and of its two neighbours are staged in shared memory once
// (a neighbour row is scanned by every slot of the frames next to it).
#define RM_FRAMES 8
#define RM_THREADS 256
__global__ void __launch_bounds__(RM_THREADS) remove_kernel(const double *__restrict__ cand_in, const double *__restrict__ score_in,
                                                            const int *__restrict__ nc_ptr, int f0_length, int max_candidates,
                                                            double *__restrict__ cand_out, double *__restrict__ score_out) {
  extern __shared__ double rm_rows[];   // [RM_FRAMES + 2][nc7]: frames first - 1 .. first + RM_FRAMES
  const int nc7 = *nc_ptr * 7;
  const int first = blockIdx.x * RM_FRAMES;
  const int n_frames = min(RM_FRAMES, f0_length - first);
  for (int e = threadIdx.x; e < (RM_FRAMES + 2) * nc7; e += RM_THREADS) {
    const int r = e / nc7, k = e - r * nc7;
    const int frame = first - 1 + r;
    // tmp_f0_candidates_ rows 0 and f0_length-1 are never filled (zero)
    const bool live = frame > 0 && frame < f0_length - 1;
    rm_rows[e] = live ? cand_in[(size_t)frame * max_candidates + k] : 0.0;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n_frames * max_candidates; e += RM_THREADS) {
    const int fi = e / max_candidates, j = e - fi * max_candidates;
    const int i = first + fi;
    const size_t g = (size_t)i * max_candidates + j;
    if (j >= nc7) {  // beyond the populated slots: nothing was ever stored there
      cand_out[g] = 0.0;
      score_out[g] = 0.0;
      continue;
    }
    double c = cand_in[g], sc = score_in[g];
    if (i >= 1 && i < f0_length - 1 && c != 0) {
      const double reference_f0 = c;
      double err[2];
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        const double *row = rm_rows + (size_t)(fi + (side == 0 ? 2 : 0)) * nc7;   // frame i + 1 / i - 1
        // min_k fl(|ref - v_k| / ref) = fl(min_k |ref - v_k| / ref): rounded division by a positive
        // constant is monotone, so one division reproduces selectBestF0's running minimum bit for bit
        double dmin = fabs(reference_f0 - row[0]);
        for (int k = 1; k < nc7; ++k) {
          const double d = fabs(reference_f0 - row[k]);
          dmin = d < dmin ? d : dmin;
        }
        const double e2 = dmin / reference_f0;
        err[side] = e2 > 1.0 ? 1.0 : e2;
      }
      const double min_error = err[0] < err[1] ? err[0] : err[1];
      if (min_error > 0.05) { c = 0; sc = 0; }
    }
    cand_out[g] = c;
    score_out[g] = sc;
  }
}

int ilog2_exact(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return ((1 << l) == n) ? l : -1;
}

int dec_pad_host(int r) { return r + (r / DEC_CHUNK); }

}  // namespace

// =============================================================================================
// host side
// =============================================================================================
int wb_harvest_plan_init(WbHarvestPlan *pl, int fs, const WbHarvestOptionInternal &opt) {
  pl->fs = fs;
  pl->opt = opt;
  // harvest.cpp:81-83
  int r = wb_round(fs / opt.target_fs);
  r = wb_max_i(wb_min_i(r, 12), 1);
  pl->decimation_ratio = r;
  pl->actual_fs = static_cast<double>(fs) / r;
  // harvest.cpp:1388-1396
  const double adjusted_f0_floor = opt.f0_floor * 0.9;
  const double adjusted_f0_ceil = opt.f0_ceil * 1.1;
  pl->nch = 1 + static_cast<int>(log(adjusted_f0_ceil / adjusted_f0_floor) / WB_LOG2 * opt.channels_in_octave);
  if (pl->nch < 2 || pl->nch > 4096) return WB_ERR_UNSUPPORTED;
  pl->boundary_f0.resize(pl->nch);
  pl->half_len.resize(pl->nch);
  int h_max = 0;
  for (int i = 0; i < pl->nch; ++i) {
    pl->boundary_f0[i] = adjusted_f0_floor * pow(2.0, static_cast<double>(i + 1) / opt.channels_in_octave);
    pl->half_len[i] = wb_round(pl->actual_fs / pl->boundary_f0[i] * 2.0);  // harvest.cpp:1264
    if (pl->half_len[i] > h_max) h_max = pl->half_len[i];
  }
  pl->h_max = h_max;
  pl->max_candidates = wb_round(pl->nch / 10) * 7;  // harvest.cpp:1418-1419 (integer division)
  if (pl->max_candidates > 128 || pl->max_candidates < 7) return WB_ERR_UNSUPPORTED;
  pl->NB = 8192;
  while (pl->NB < 8 * (h_max + 1) && pl->NB < 16384) pl->NB *= 2;
  if (2 * h_max + 2 >= pl->NB / 2) return WB_ERR_UNSUPPORTED;
  pl->V = pl->NB - 2 - 2 * h_max;
  if (r > 1 && !decimate_coefficients(r, (DecimCoef *)pl->decim_coef)) return WB_ERR_UNSUPPORTED;
  pl->filters_ready = false;
  if (!pl->aux_stream) {
    if (cudaStreamCreateWithFlags(&pl->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&pl->aux_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&pl->aux_join, cudaEventDisableTiming) != cudaSuccess)
      return WB_ERR_CUDA;
  }
  return WB_OK;
}

static int harvest_prepare_filters(WbHarvestPlan *pl, WbWorkspace *ws, cudaStream_t stream) {
  if (pl->filters_ready) return WB_OK;
  const int NC = pl->NB / 2;
  double *d_bf = (double *)ws->get("hv_bf", sizeof(double) * pl->nch);
  int *d_hl = (int *)ws->get("hv_hl", sizeof(int) * pl->nch);
  cplx *d_Hc = (cplx *)ws->get("hv_Hc", sizeof(cplx) * (size_t)pl->nch * (NC + 1));
  if (!d_bf || !d_hl || !d_Hc) return WB_ERR_CUDA;
  WB_CUDA_CHECK(cudaMemcpyAsync(d_bf, pl->boundary_f0.data(), sizeof(double) * pl->nch, cudaMemcpyHostToDevice, stream));
  WB_CUDA_CHECK(cudaMemcpyAsync(d_hl, pl->half_len.data(), sizeof(int) * pl->nch, cudaMemcpyHostToDevice, stream));
  const cplx *tw = wb_twiddle_table(pl->NB);
  if (!tw) return WB_ERR_CUDA;
  const size_t smem = sizeof(cplx) * wb_fft_slots(NC);
  int rc = WB_DISPATCH_LOG2(ilog2_exact(pl->NB), 13, 14, {
    if (cudaFuncSetAttribute(filter_spec_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WB_ERR_CUDA;
    WB_LAUNCH("filter_spec_kernel", filter_spec_kernel<L2><<<pl->nch, 512, smem, stream>>>(d_bf, d_hl, pl->actual_fs, tw, d_Hc));
  });
  if (rc) return rc;
  WB_CUDA_CHECK(cudaGetLastError());
  WB_CUDA_CHECK(cudaStreamSynchronize(stream));  // host vectors are pageable; make the plan self-contained
  pl->filters_ready = true;
  return WB_OK;
}

// Runs Harvest at a 1 ms (frame_period = 1) grid up to the pruned candidate table and then the
// tail; writes d_f0_basic[Lb].
int wb_harvest_run_basic(WbHarvestPlan *pl, WbWorkspace *ws, const double *d_x, int x_length, int frame_period,
                         double *d_f0_basic, int *f0_length_out, cudaStream_t stream) {
  const int fs = pl->fs, r = pl->decimation_ratio;
  const double afs = pl->actual_fs;
  int rc;
  if ((rc = harvest_prepare_filters(pl, ws, stream))) return rc;
  const int NB = pl->NB, NC = NB / 2, log2nc = ilog2_exact(NC);
  const int y_length = 1 + static_cast<int>(x_length / r);                               // harvest.cpp:1399
  const int Lb = static_cast<int>(1000.0 * x_length / fs / frame_period) + 1;            // harvest.cpp:173-176
  *f0_length_out = Lb;
  const int nch = pl->nch, MC = pl->max_candidates, own_cap = MC / 7;
  // one call handles up to IV_MAX_BLOCKS overlap-save blocks (~15 min at the 8 kHz analysis rate): checked before
  // anything is enqueued; longer streams go through the segmenting path (worldb200/parallel.py)
  if ((y_length + pl->V - 1) / pl->V > IV_MAX_BLOCKS) return WB_ERR_UNSUPPORTED;

  // ---- clears of the candidate tables: nothing before the candidate stage touches them, so they run on
  // the auxiliary stream beside the chain
  {
    int *d_nc0 = (int *)ws->get("hv_nc", 16);
    double *d_candA0 = (double *)ws->get("hv_candA", sizeof(double) * (size_t)Lb * MC * 2);
    if (!d_nc0 || !d_candA0) return WB_ERR_CUDA;
    cudaStream_t aux = wb_prof_is_enabled() ? stream : pl->aux_stream;
    WB_CUDA_CHECK(cudaEventRecord(pl->aux_fork, stream));
    WB_CUDA_CHECK(cudaStreamWaitEvent(aux, pl->aux_fork, 0));
    WB_CUDA_CHECK(cudaMemsetAsync(d_nc0, 0, 16, aux));  // [0] = nc, [1] = number of own candidates
    WB_CUDA_CHECK(cudaMemsetAsync(d_candA0, 0, sizeof(double) * (size_t)Lb * MC * 2, aux));
    WB_CUDA_CHECK(cudaEventRecord(pl->aux_join, aux));
  }

  // ---- H1/H2: decimated, DC-"corrected" waveform
  double *d_y = (double *)ws->get("hv_y", sizeof(double) * (y_length + 8));
  unsigned long long *d_absmax = (unsigned long long *)ws->get("hv_absmax", 16);
  if (!d_y || !d_absmax) return WB_ERR_CUDA;
  WB_CUDA_CHECK(cudaMemsetAsync(d_absmax, 0, 8, stream));
  if (r == 1) {
    WB_LAUNCH("copy_kernel", copy_kernel<<<(x_length + 255) / 256, 256, 0, stream>>>(d_x, x_length, d_y, d_absmax));  // y_length = x_length + 1: last is zero
    WB_CUDA_CHECK(cudaMemsetAsync(d_y + x_length, 0, sizeof(double) * (y_length - x_length), stream));
  } else {
    const int lag = static_cast<int>(ceil(140.0 / r) * r);                               // harvest.cpp:222
    const int len1 = x_length + lag * 2;
    const int len2 = len1 + 2 * DEC_NFACT;
    double *d_fwd = (double *)ws->get("hv_fwd", sizeof(double) * len2);
    if (!d_fwd) return WB_ERR_CUDA;
    DecimCoef dc;
    memcpy(&dc, pl->decim_coef, sizeof(dc));
    const int n_tiles = (len2 + DEC_TILE - 1) / DEC_TILE;
    const size_t dec_smem_f = sizeof(double) * (dec_pad_host(DEC_TILE + DEC_WARM) + 1 + dec_pad_host(DEC_TILE) + 1);
    const size_t dec_smem_b = sizeof(double) * (dec_pad_host(DEC_TILE + DEC_WARM) + 1);
    WB_CUDA_CHECK(cudaFuncSetAttribute(dec_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_smem_f));
    WB_CUDA_CHECK(cudaFuncSetAttribute(dec_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_smem_b));
    WB_CUDA_CHECK(cudaMemsetAsync(d_y, 0, sizeof(double) * y_length, stream));            // new_y is zero-initialised
    WB_LAUNCH("dec_forward_kernel", dec_forward_kernel<<<n_tiles, DEC_BLOCK, dec_smem_f, stream>>>(d_x, x_length, lag, len1, len2, dc, d_fwd));
    WB_LAUNCH("dec_backward_kernel", dec_backward_kernel<<<n_tiles, DEC_BLOCK, dec_smem_b, stream>>>(d_fwd, len1, len2, r, lag, dc, y_length, d_y, d_absmax));
  }
  WB_LAUNCH("dc_fix_kernel", dc_fix_kernel<<<1, 1024, 0, stream>>>(d_y, y_length, d_absmax));
  WB_CUDA_CHECK(cudaGetLastError());

  // ---- H3: overlap-save block spectra
  const int n_blocks = (y_length + pl->V - 1) / pl->V;
  cplx *d_Yb = (cplx *)ws->get("hv_Yb", sizeof(cplx) * (size_t)n_blocks * (NC + 1));
  if (!d_Yb) return WB_ERR_CUDA;
  const cplx *tw = wb_twiddle_table(NB);
  const size_t smem_fft = sizeof(cplx) * wb_fft_slots(NC);
  rc = WB_DISPATCH_LOG2(log2nc + 1, 13, 14, {
    if (cudaFuncSetAttribute(yspec_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fft) != cudaSuccess) return WB_ERR_CUDA;
    WB_LAUNCH("yspec_kernel", yspec_kernel<L2><<<n_blocks, 512, smem_fft, stream>>>(d_y, y_length, pl->V, pl->h_max, tw, d_Yb));
  });
  if (rc) return rc;
  WB_CUDA_CHECK(cudaGetLastError());

  // ---- H5: channels
  const int ecap = y_length / 2 + 4;
  const int bcap = pl->V / 2 + 2;
  if (n_blocks > IV_MAX_BLOCKS) return WB_ERR_UNSUPPORTED;  // ~15 min; longer streams are processed in segments (parallel.py)
  int *d_ecount = (int *)ws->get("hv_ecount", sizeof(int) * nch * 4);
  double *d_seg = (double *)ws->get("hv_seg_edges", sizeof(double) * (size_t)nch * 4 * n_blocks * bcap);
  int *d_segc = (int *)ws->get("hv_seg_count", sizeof(int) * (size_t)nch * 4 * n_blocks);
  if (!d_ecount || !d_seg || !d_segc) return WB_ERR_CUDA;
  {
    ChanParams p;
    p.Yb = d_Yb; p.Hc = (const cplx *)ws->find("hv_Hc"); p.half_len = (const int *)ws->find("hv_hl");
    p.n_blocks = n_blocks; p.NB = NB; p.log2nc = log2nc; p.V = pl->V; p.h_max = pl->h_max; p.y_length = y_length;
    p.tw = tw; p.seg_edges = d_seg; p.seg_count = d_segc; p.bcap = bcap;
    dim3 grid(n_blocks, nch);
    rc = WB_DISPATCH_LOG2(log2nc + 1, 13, 14, {
      if (cudaFuncSetAttribute(channel_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fft) != cudaSuccess) return WB_ERR_CUDA;
      WB_LAUNCH("channel_kernel", channel_kernel<L2><<<grid, CH_THREADS, smem_fft, stream>>>(p));
    });
    if (rc) return rc;
    WB_CUDA_CHECK(cudaGetLastError());
    WB_CUDA_CHECK(cudaGetLastError());
  }

  // ---- interval sequences -> the four interpolated contours on the frame grid
  double *d_contour = (double *)ws->get("hv_contour", sizeof(double) * (size_t)nch * 4 * ((Lb + 31) / 32) * 32);
  if (!d_contour) return WB_ERR_CUDA;
  {
    IntervalParams p;
    p.seg_edges = d_seg; p.seg_count = d_segc; p.n_blocks = n_blocks; p.bcap = bcap;
    p.ecount = d_ecount; p.ecap = ecap; p.fs = afs;
    p.f0_length = Lb; p.frame_period = frame_period; p.contour = d_contour;
    // frame times: a function of (Lb, frame_period) only, tabulated once per plan and length
    double *d_ttab = (double *)ws->get("hv_ttab", sizeof(double) * Lb);
    if (!d_ttab) return WB_ERR_CUDA;
    if (pl->ttab_len != Lb || pl->ttab_period != frame_period || pl->ttab_ptr != (const void *)d_ttab) {
      WB_LAUNCH("frame_time_kernel", frame_time_kernel<<<(Lb + 255) / 256, 256, 0, stream>>>(Lb, frame_period, d_ttab));
      pl->ttab_len = Lb; pl->ttab_period = frame_period; pl->ttab_ptr = d_ttab;
    }
    p.t_tab = d_ttab;
    WB_LAUNCH("interval_kernel", interval_kernel<<<nch * 4, IV_THREADS, 0, stream>>>(p));
    WB_CUDA_CHECK(cudaGetLastError());
  }

  // ---- raw candidates, official candidates, refinement work list
  double *d_raw = (double *)ws->get("hv_raw", sizeof(double) * (size_t)nch * Lb);
  double *d_own = (double *)ws->get("hv_own", sizeof(double) * (size_t)Lb * own_cap);
  int *d_nc = (int *)ws->get("hv_nc", 16);
  int *d_work = (int *)ws->get("hv_work", sizeof(int) * (size_t)Lb * own_cap);
  double *d_candA = (double *)ws->get("hv_candA", sizeof(double) * (size_t)Lb * MC * 2);   // cand | score, one memset
  double *d_candB = (double *)ws->get("hv_candB", sizeof(double) * (size_t)Lb * MC);
  double *d_scoreB = (double *)ws->get("hv_scoreB", sizeof(double) * (size_t)Lb * MC);
  if (!d_raw || !d_own || !d_nc || !d_work || !d_candA || !d_candB || !d_scoreB) return WB_ERR_CUDA;
  double *d_scoreA = d_candA + (size_t)Lb * MC;
  if (own_cap > 32) return WB_ERR_UNSUPPORTED;
  WB_CUDA_CHECK(cudaStreamWaitEvent(stream, pl->aux_join, 0));   // the table clears issued at the top
  {
    CandParams p;
    p.contour = d_contour; p.ecount = d_ecount; p.boundary_f0 = (const double *)ws->find("hv_bf");
    p.nch = nch; p.f0_length = Lb; p.f0_floor = pl->opt.f0_floor; p.f0_ceil = pl->opt.f0_ceil;
    p.raw = d_raw; p.own = d_own; p.own_cap = own_cap; p.nc_max = d_nc; p.work = d_work;
    const size_t smem = sizeof(double) * (size_t)nch * CD_PITCH;
    if (smem > 200 * 1024) return WB_ERR_UNSUPPORTED;
    WB_CUDA_CHECK(cudaFuncSetAttribute(candidate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WB_LAUNCH("candidate_kernel", candidate_kernel<<<(Lb + CD_FRAMES - 1) / CD_FRAMES, CD_THREADS, smem, stream>>>(p));
    WB_CUDA_CHECK(cudaGetLastError());
  }
  const long long n_cs = (long long)Lb * MC;
  {
    RefineParams p;
    p.y = d_y; p.y_length = y_length; p.actual_fs = afs; p.f0_floor = pl->opt.f0_floor; p.f0_ceil = pl->opt.f0_ceil;
    p.frame_period = frame_period; p.work = d_work; p.nc_and_count = d_nc; p.max_candidates = MC;
    p.own = d_own; p.own_cap = own_cap; p.f0_length = Lb;
    p.cand = d_candA; p.score = d_scoreA;
    // candidates are in [f0_floor, f0_ceil] of the raw stage: half window <= 1.5 fs / (0.9 floor) + 1
    const int max_hw = static_cast<int>(1.5 * afs / (pl->opt.f0_floor * 0.9) + 1.0) + 1;
    p.max_wlen = 2 * max_hw + 1;
    const int max_log2 = 2 + (int)floor(log2((double)p.max_wlen));
    if (max_log2 > 15) return WB_ERR_UNSUPPORTED;
    for (int l = 0; l < 16; ++l) p.tw[l] = nullptr;
    for (int l = 3; l <= max_log2; ++l) {
      p.tw[l] = wb_twiddle_table(1 << l);
      if (!p.tw[l]) return WB_ERR_CUDA;
    }
    // per-candidate constants by one thread per candidate, then the accumulation by one warp per candidate
    RefineSetup *d_setup = (RefineSetup *)ws->get("hv_refine_setup", sizeof(RefineSetup) * (size_t)Lb * own_cap);
    if (!d_setup) return WB_ERR_CUDA;
    WB_LAUNCH("refine_setup_kernel", refine_setup_kernel<<<wb_sm_count() * 8, 128, 0, stream>>>(p, d_setup));
    WB_LAUNCH("refine_kernel", refine_mma_kernel<<<wb_sm_count() * 8, RF_WARPS * 32, 0, stream>>>(p, d_setup, d_nc + 2));
    WB_CUDA_CHECK(cudaGetLastError());
  }
  (void)n_cs;
  WB_LAUNCH("remove_kernel", remove_kernel<<<(Lb + RM_FRAMES - 1) / RM_FRAMES, RM_THREADS, sizeof(double) * (RM_FRAMES + 2) * MC, stream>>>(d_candA, d_scoreA, d_nc, Lb, MC, d_candB, d_scoreB));
  WB_CUDA_CHECK(cudaGetLastError());

  // ---- contour fixing + smoothing
  return wb_harvest_tail(ws, d_candB, d_scoreB, d_nc, Lb, MC, d_f0_basic, stream);
}

// ---- decimate() as a stand-alone call (include/world_matlabfunctions.hpp; world_matlabfunctions.cpp:184-210) ----
extern "C" int wb_decimate_length(int x_length, int r) {
  if (x_length <= 0 || r < 1) return 0;
  const int nout = x_length / r + 1;
  const int nbeg = r - r * nout + x_length;
  return (x_length + DEC_NFACT - nbeg + r - 1) / r;
}

extern "C" int wb_decimate(const double *x, int x_length, int r, double *y) {
  if (!x || !y || x_length < 2 * DEC_NFACT) return WB_ERR_ARG;
  DecimCoef dc;
  if (!decimate_coefficients(r, &dc)) return WB_ERR_UNSUPPORTED;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return WB_ERR_CUDA;
  const int y_length = wb_decimate_length(x_length, r);
  const int len1 = x_length, len2 = len1 + 2 * DEC_NFACT;
  double *d_x = nullptr, *d_fwd = nullptr, *d_y = nullptr;
  unsigned long long *d_absmax = nullptr;
  cudaStream_t stream = nullptr;
  int rc = WB_OK;
  if (cudaMalloc(&d_x, sizeof(double) * x_length) != cudaSuccess || cudaMalloc(&d_fwd, sizeof(double) * len2) != cudaSuccess ||
      cudaMalloc(&d_y, sizeof(double) * y_length) != cudaSuccess || cudaMalloc(&d_absmax, 16) != cudaSuccess)
    rc = WB_ERR_CUDA;
  if (!rc) {
    const int n_tiles = (len2 + DEC_TILE - 1) / DEC_TILE;
    const size_t smem_f = sizeof(double) * (dec_pad_host(DEC_TILE + DEC_WARM) + 1 + dec_pad_host(DEC_TILE) + 1);
    const size_t smem_b = sizeof(double) * (dec_pad_host(DEC_TILE + DEC_WARM) + 1);
    cudaError_t e = cudaMemcpy(d_x, x, sizeof(double) * x_length, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(d_absmax, 0, 16);
    if (e == cudaSuccess) e = cudaMemset(d_y, 0, sizeof(double) * y_length);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dec_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dec_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
    if (e == cudaSuccess) {
      WB_LAUNCH("dec_forward_kernel", dec_forward_kernel<<<n_tiles, DEC_BLOCK, smem_f, stream>>>(d_x, x_length, 0, len1, len2, dc, d_fwd));
      WB_LAUNCH("dec_backward_kernel", dec_backward_kernel<<<n_tiles, DEC_BLOCK, smem_b, stream>>>(d_fwd, len1, len2, r, 0, dc, y_length, d_y, d_absmax));
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(y, d_y, sizeof(double) * y_length, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = WB_ERR_CUDA;
  }
  cudaFree(d_x); cudaFree(d_fwd); cudaFree(d_y); cudaFree(d_absmax);
  return rc;
}
