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
#include "wb_smooth.cuh"

#include <math.h>
#include <stdlib.h>
#include <vector>

namespace {

#define D4C_HANNING 1
#define D4C_BLACKMAN 2
#define D4C_MAX_AP 8
#define D4C_BODY_THREADS 256

__device__ __forceinline__ int d4c_half_window(double ratio, int fs, double f0) {
  return wb_round(ratio * fs / f0 / 2.0);  // d4c.cpp:250
}

// d4c.cpp:246-303.  at(j) -> reference to destination sample j; win(j) -> reference to scratch for
// window sample j.  All threads must call; ends with __syncthreads().  Returns the window length.
template <typename At, typename Win>
__device__ inline int d4c_windowed_waveform(const double *__restrict__ x, int x_length, int fs, double f0,
                                            double position_s, int window_type, double ratio,
                                            const double *__restrict__ noise, Win win, double *red, At at,
                                            const double *rot = nullptr) {
  const int hw = d4c_half_window(ratio, fs, f0);
  const int wlen = 2 * hw + 1;
  const int origin = wb_round(position_s * fs + 0.001);
  const double c1 = 2.0 / ratio / fs;
  const double c2 = WB_PI * f0;
  // The window angle c2 * c1 * (j - hw) is linear in j: one sincos at the thread's first sample (the
  // reference's expression), then a rotation by blockDim samples per step (<= 32 steps, ~1e-15 drift);
  // cos(2t) = 2 cos(t)^2 - 1 for the Blackman term (d4c.cpp:266-283).
  // (`rot`: the four rotation constants, if the caller already has them for this f0 and ratio)
  double sd, cd, sn, cs;
  if (rot) {
    sd = rot[0]; cd = rot[1]; sn = rot[2]; cs = rot[3];
  } else {
    sincos(c2 * c1 * blockDim.x, &sd, &cd);
    sincos(c2 * (c1 * ((int)threadIdx.x - hw)), &sn, &cs);
  }
  double s1 = 0.0, s2 = 0.0;
  // (the waveform and noise samples of the NEXT step are requested before this step's arithmetic: the loop used
  // to sit on these two loads)
  double x_next = 0.0, n_next = 0.0;
  if ((int)threadIdx.x < wlen) {
    x_next = x[wb_min_i(x_length - 1, wb_max_i(0, origin + (int)threadIdx.x - hw))];
    n_next = noise[threadIdx.x];
  }
  for (int j = threadIdx.x; j < wlen; j += blockDim.x) {
    const double x_here = x_next, n_here = n_next;
    const int jn = j + blockDim.x;
    if (jn < wlen) {
      x_next = x[wb_min_i(x_length - 1, wb_max_i(0, origin + jn - hw))];
      n_next = noise[jn];
    }
    double w;
    if (window_type == D4C_HANNING) w = 0.5 * cs + 0.5;
    else w = 0.42 + 0.5 * cs + 0.08 * (2.0 * cs * cs - 1.0);
    win(j) = w;
    const double v = x_here * w + n_here * WB_SAFEGUARD;
    at(j) = v;
    s1 += v;
    s2 += w;
    const double c_next = cs * cd - sn * sd;
    sn = sn * cd + cs * sd;
    cs = c_next;
  }
  wb_block_sum2(s1, s2, red);
  const double coef = s1 / s2;
  for (int j = threadIdx.x; j < wlen; j += blockDim.x) at(j) -= win(j) * coef;
  __syncthreads();
  return wlen;
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

// ---- Love Train (d4c.cpp:181-240) ---------------------------------------------------------
struct LtParams {
  const double *x; int x_length;
  const double *tpos; const double *f0; int f0_length;
  int fs; int fft_size; int log2nc; double lowest_f0;
  int boundary0, boundary1, boundary2;
  const cplx *twiddle;
  const double *noise; const unsigned long long *noise_off;
  double *ap0;
  int frame_begin;   // this launch covers frames frame_begin + blockIdx.x
};

template <int LOG2N>
__global__ void __launch_bounds__(256) lt_frame_kernel(LtParams p) {
  extern __shared__ double2 smem_raw[];
  constexpr int N = 1 << LOG2N, NC = N / 2;
  cplx *S = smem_raw;
  double *win = reinterpret_cast<double *>(S + wb_fft_slots(NC));  // N doubles
  double *red = win + N;                                          // 128
  double *W = reinterpret_cast<double *>(S);
  const int frame = p.frame_begin + blockIdx.x;
  const double f0 = p.f0[frame];
  if (f0 == 0.0) {
    if (threadIdx.x == 0) p.ap0[frame] = 0.0;
    return;
  }
  const double cf0 = f0 > p.lowest_f0 ? f0 : p.lowest_f0;
  const int wlen = d4c_windowed_waveform(p.x, p.x_length, p.fs, cf0, p.tpos[frame], D4C_BLACKMAN, 3.0,
                                         p.noise + p.noise_off[frame], [&](int j) -> double & { return win[j]; }, red,
                                         [&](int j) -> double & { return W[wb_didx(j)]; });
  for (int j = wlen + threadIdx.x; j < N; j += blockDim.x) W[wb_didx(j)] = 0.0;
  __syncthreads();
  // power in (boundary0, boundary1] and (boundary0, boundary2]; bins above N/2 count as zero
  double a = 0.0, b = 0.0;
  const int b0 = p.boundary0, b1 = p.boundary1, b2 = p.boundary2;
  // wb_rfft's emit runs once per k on some thread: accumulate per thread, reduce afterwards
  wb_rfft_t<1, LOG2N - 1>(S, p.twiddle, [&](int k, cplx X) {
    if (k > b0 && k <= b2) {
      const double pw = X.x * X.x + X.y * X.y;
      b += pw;
      if (k <= b1) a += pw;
    }
  });
  wb_block_sum2(a, b, red);
  if (threadIdx.x == 0) p.ap0[frame] = a / b;
}

// ---- order statistic: sum of the m smallest of v[0..n) (d4c.cpp:494-499) ----------------------
// The reference sorts the band power spectrum and takes a cumulative sum; only
// S[bins - boundary - 2] / S[bins - 1] is used, i.e. (sum of the m smallest) / total.  Radix select
// on the bit patterns of the (non-negative) values, two arrays at once (the two bands of one paired
// transform).
//  * Keys are the bit patterns minus SEL_BIAS (the pattern of 2^-255, saturating at 0): every value in
//    [2^-255, 2^257) then has its three top key bits clear, and the bits just below them -- nine exponent
//    bits and the leading mantissa bits -- form a first digit that does not straddle the exponent
//    boundary at 2.0 the way the raw patterns of a power spectrum (1e-10 .. 1e7) do.
//  * The PRODUCER of the values counts that first digit (PRE_BITS wide, packed 16-bit counters) while it
//    computes them; if all keys are in range the select starts from the finished histogram and needs no
//    counting pass over the data.  Otherwise, and for further digits, counting passes with BITS-wide
//    digits start at the first key bit that actually varies (AND / OR of the keys, also gathered by
//    the producer).
//  * As soon as the bucket holding the m-th smallest key has <= SEL_LIST keys, ONE pass sums everything
//    below the bucket and collects the bucket's keys, and a single warp ranks them.
// get(i, w) -> value i of array w.  hist: 2 << BITS ints (= 2 << PRE_BITS shorts); ctl: SEL_CTL_WORDS +
// 2 * SEL_LIST words.
#define SEL_LIST 32
#define SEL_CTL_WORDS (12 + D4C_BODY_THREADS / 32)
#define SEL_BIAS 0x3000000000000000ull
#define SEL_PRE_TOP 61   // the pre-counted digit ends below key bit 61
__device__ __forceinline__ unsigned long long sel_key(double v) {
  const unsigned long long k = (unsigned long long)__double_as_longlong(v);
  return k > SEL_BIAS ? k - SEL_BIAS : 0ull;
}
__device__ __forceinline__ double sel_val(unsigned long long key) { return __longlong_as_double((long long)(key + SEL_BIAS)); }
// producer side: count key `k` of array w (0 / 1) in the packed histogram
template <int PRE_BITS>
__device__ __forceinline__ void sel_precount(unsigned *hist16, int w, unsigned long long k) {
  const int bin = (int)((k >> (SEL_PRE_TOP - PRE_BITS)) & ((1ull << PRE_BITS) - 1ull));
  const int at = (w << PRE_BITS) + bin;
  atomicAdd(&hist16[at >> 1], 1u << (16 * (at & 1)));
}

template <int BITS, int PRE_BITS, typename Get>
__device__ inline void d4c_sum_smallest2(Get get, int n, int m, bool has_second, unsigned long long and_a,
                                         unsigned long long or_a, unsigned long long and_b, unsigned long long or_b,
                                         bool precounted, int *hist, unsigned long long *ctl, double *red,
                                         double &low_a, double &low_b) {
  constexpr int BINS = 1 << BITS;
  constexpr int PER = BINS >= D4C_BODY_THREADS ? BINS / D4C_BODY_THREADS : 1;  // bins per thread in the scan
  constexpr int PRE_BINS = 1 << PRE_BITS;
  constexpr int PRE_PER = PRE_BINS >= D4C_BODY_THREADS ? PRE_BINS / D4C_BODY_THREADS : 1;
  constexpr int PRE_SHIFT = SEL_PRE_TOP - PRE_BITS;
  constexpr int NW = D4C_BODY_THREADS / 32;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  int *s_wtot = reinterpret_cast<int *>(ctl + 12);  // 2 x NW warp totals of the bucket scan (NW words)
  unsigned long long *list = ctl + SEL_CTL_WORDS;    // 2 x SEL_LIST keys
  // ---- AND / OR of the keys of each array
  {
    const unsigned full = 0xffffffffu;
    const unsigned aah = __reduce_and_sync(full, (unsigned)(and_a >> 32)), aal = __reduce_and_sync(full, (unsigned)and_a);
    const unsigned oah = __reduce_or_sync(full, (unsigned)(or_a >> 32)), oal = __reduce_or_sync(full, (unsigned)or_a);
    const unsigned abh = __reduce_and_sync(full, (unsigned)(and_b >> 32)), abl = __reduce_and_sync(full, (unsigned)and_b);
    const unsigned obh = __reduce_or_sync(full, (unsigned)(or_b >> 32)), obl = __reduce_or_sync(full, (unsigned)or_b);
    if (tid == 0) {
      ctl[8] = ~0ull; ctl[9] = 0ull; ctl[10] = ~0ull; ctl[11] = 0ull;
      ctl[1] = (unsigned long long)m; ctl[5] = (unsigned long long)m;   // remaining rank (1-based)
    }
    __syncthreads();
    if (lane == 0) {
      atomicAnd(&ctl[8], ((unsigned long long)aah << 32) | aal);
      atomicOr(&ctl[9], ((unsigned long long)oah << 32) | oal);
      atomicAnd(&ctl[10], ((unsigned long long)abh << 32) | abl);
      atomicOr(&ctl[11], ((unsigned long long)obh << 32) | obl);
    }
    __syncthreads();
  }
  unsigned long long prefix[2], mask[2];
  int shift[2], width[2];
  bool done[2];       // threshold fully determined (= prefix)
  bool listed[2];     // bucket small enough: resolve from the collected list
  bool counted[2];    // the first digit of this array is already in the packed histogram
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    const unsigned long long all_and = ctl[8 + 2 * w], all_or = ctl[9 + 2 * w];
    const unsigned long long diff = all_and ^ all_or;
    listed[w] = false;
    counted[w] = false;
    if (diff == 0ull || (w == 1 && !has_second)) {  // every key identical: that key is the threshold
      prefix[w] = all_and; mask[w] = ~0ull; shift[w] = 0; width[w] = 0; done[w] = true;
    } else if (precounted && (all_or >> SEL_PRE_TOP) == 0ull) {
      prefix[w] = 0ull; mask[w] = ~((1ull << SEL_PRE_TOP) - 1ull);   // bits 61.. are zero in every key
      shift[w] = PRE_SHIFT; width[w] = PRE_BITS; done[w] = false; counted[w] = true;
    } else {
      const int top = 64 - __clzll((long long)diff);  // bits [0, top) vary
      mask[w] = (top >= 64) ? 0ull : ~((1ull << top) - 1ull);
      prefix[w] = all_and & mask[w];
      shift[w] = top > BITS ? top - BITS : 0;
      width[w] = top - shift[w];
      done[w] = false;
    }
  }
  // ---- bucket holding the remaining rank: block-wide scan of the counters.  count(w, bin) reads a counter.
  auto find_bucket = [&](auto count, int per, int nbins, bool act0, bool act1) {
    int c[2] = {0, 0}, incl[2];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      if (w == 0 ? act0 : act1)
        for (int q = 0; q < per; ++q) c[w] += (tid * per + q < nbins) ? count(w, tid * per + q) : 0;
      incl[w] = c[w];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl[w], o);
        if (lane >= o) incl[w] += t;
      }
      if (lane == 31) s_wtot[w * NW + warp] = incl[w];
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      if (!(w == 0 ? act0 : act1)) continue;
      int before = 0;
      for (int q = 0; q < warp; ++q) before += s_wtot[w * NW + q];
      const int excl = before + incl[w] - c[w];
      unsigned long long *c4 = ctl + w * 4;
      const int remaining = (int)c4[1];
      if (remaining > excl && remaining <= excl + c[w]) {
        int r = remaining - excl;
        int d = tid * per, hcount = 0;
        for (int q = 0; q < per && tid * per + q < nbins; ++q) {
          const int hv = count(w, tid * per + q);
          if (r <= hv) { d = tid * per + q; hcount = hv; break; }
          r -= hv;
        }
        c4[0] = (unsigned long long)d;
        c4[2] = (unsigned long long)r;
        c4[3] = (unsigned long long)hcount;
      }
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      if (w == 0 ? act0 : act1) {
        prefix[w] |= ctl[w * 4 + 0] << shift[w];
        mask[w] |= ((1ull << width[w]) - 1ull) << shift[w];
        if (shift[w] == 0) done[w] = true;
        else if (ctl[w * 4 + 3] <= (unsigned long long)SEL_LIST) listed[w] = true;
        const int ns = shift[w] > BITS ? shift[w] - BITS : 0;
        width[w] = shift[w] - ns;
        shift[w] = ns;
      }
    }
    __syncthreads();
    if (tid == 0) { ctl[1] = ctl[2]; ctl[5] = ctl[6]; }
  };
  if (counted[0] || counted[1]) {
    const unsigned *h16 = reinterpret_cast<const unsigned *>(hist);
    find_bucket([&](int w, int bin) { const int at = (w << PRE_BITS) + bin; return (int)((h16[at >> 1] >> (16 * (at & 1))) & 0xffffu); },
                PRE_PER, PRE_BINS, counted[0], counted[1]);
  }
  for (int pass = 0; pass < 64; ++pass) {
    const bool act0 = !(done[0] || listed[0]), act1 = !(done[1] || listed[1]);
    if (!act0 && !act1) break;
    {
      int4 *h4 = reinterpret_cast<int4 *>(hist);
      for (int i = tid; i < 2 * BINS / 4; i += nt)
        if ((i < BINS / 4) ? act0 : act1) h4[i] = make_int4(0, 0, 0, 0);
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        if (w == 0 ? act0 : act1) {
          const unsigned long long key = sel_key(get(i, w));
          if ((key & mask[w]) == prefix[w])
            atomicAdd(&hist[w * BINS + (int)((key >> shift[w]) & ((1ull << width[w]) - 1ull))], 1);
        }
      }
    }
    __syncthreads();
    find_bucket([&](int w, int bin) { return hist[w * BINS + bin]; }, PER, BINS, act0, act1);
  }
  // ---- one pass: sum / count of everything below the bucket, collect the bucket's keys
  if (tid == 0) { s_wtot[0] = 0; s_wtot[1] = 0; }
  __syncthreads();
  double sa = 0.0, ca = 0.0, sb = 0.0, cb = 0.0;
  for (int i = tid; i < n; i += nt) {
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      if (w == 1 && !has_second) continue;
      const double v = get(i, w);
      const unsigned long long key = sel_key(v);
      const unsigned long long top = key & mask[w];
      if (top < prefix[w]) {
        if (w == 0) { sa += v; ca += 1.0; } else { sb += v; cb += 1.0; }
      } else if (listed[w] && top == prefix[w]) {
        const int at = atomicAdd(&s_wtot[w], 1);
        if (at < SEL_LIST) list[w * SEL_LIST + at] = key;
      }
    }
  }
  __syncthreads();
  // ---- rank the collected keys: warp w resolves array w (remaining rank r, 1-based, inside the bucket)
  if ((warp == 0 && listed[0]) || (warp == 1 && listed[1])) {
    const int w = warp;
    const int cnt = min(s_wtot[w], SEL_LIST);
    const int r = (int)ctl[w * 4 + 1];
    const unsigned long long mine = lane < cnt ? list[w * SEL_LIST + lane] : ~0ull;
    int rank = 0;
    for (int j = 0; j < cnt; ++j) {
      const unsigned long long other = __shfl_sync(0xffffffffu, mine, j);
      rank += (other < mine || (other == mine && j < lane)) ? 1 : 0;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, lane < cnt && rank == r - 1);
    const int src = hit ? (__ffs(hit) - 1) : 0;
    const unsigned long long thr = __shfl_sync(0xffffffffu, mine, src);
    // bucket keys strictly below the threshold join the "below" totals; they are summed in rank order
    // (the collection order above depends on the atomics and must not leak into the rounding)
    __syncwarp();
    if (lane < cnt) list[w * SEL_LIST + rank] = mine;
    __syncwarp();
    const unsigned long long sorted = lane < cnt ? list[w * SEL_LIST + lane] : ~0ull;
    double ls = 0.0, lc = 0.0;
    if (lane < cnt && sorted < thr) { ls = sel_val(sorted); lc = 1.0; }
    ls = wb_warp_sum(ls);
    lc = wb_warp_sum(lc);
    if (lane == 0) {
      ctl[w * 4 + 0] = thr;
      ctl[w * 4 + 2] = (unsigned long long)__double_as_longlong(ls);
      ctl[w * 4 + 3] = (unsigned long long)__double_as_longlong(lc);
    }
  }
  wb_block_sum2(sa, ca, red);   // (contains the barriers that publish ctl)
  wb_block_sum2(sb, cb, red);
  double t[2];
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    t[w] = sel_val(prefix[w]);
    if (listed[w]) {
      t[w] = sel_val(ctl[w * 4 + 0]);
      const double ls = __longlong_as_double((long long)ctl[w * 4 + 2]), lc = __longlong_as_double((long long)ctl[w * 4 + 3]);
      if (w == 0) { sa += ls; ca += lc; } else { sb += ls; cb += lc; }
    }
  }
  low_a = sa + (m - ca) * t[0];
  low_b = sb + (m - cb) * t[1];
}

// ---- body (d4c.cpp:308-503 + :155-168) -------------------------------------------------------
struct BodyParams {
  const double *x; int x_length;
  const double *tpos; const double *f0; const double *ap0; int f0_length;
  int fs; int fft_size_d4c; int log2n; double threshold;
  int n_ap; int window_length;         // Nuttall window of the band analysis
  const double *nuttall;               // device, window_length doubles
  const cplx *tw_n;                    // fft_size_d4c entries (real transforms)
  const cplx *tw_2n;                   // 2*fft_size_d4c entries (complex transform of size N)
  const double *noise; const unsigned long long *noise_off;
  int out_fft_size;                    // bins of the output rows
  double *ap;                          // [f0_length][out_fft_size/2+1]
  int seg_capacity;
  int *error_flag;
  int frame_begin;                     // this launch covers frames frame_begin + blockIdx.x
  int debug_skip;                      // profiling only (WB_D4C_SKIP, profiles/d4c_phases.py): phases to leave out
};


// Shared memory (N = 4096: 110 KB -> two CTAs per SM):
//   S   : slots of an N-point complex FFT; doubles as the linear-smoothing scratch (`seg`) and,
//         in its upper half, as window scratch while only a packed real transform lives in it
//   SC  : static centroid -> static group delay
//   SP  : smoothed power  -> second smoothing output -> band power spectrum
// WARP_BANDS (EXPERIMENTAL, off unless WB_D4C_WARP_BANDS=1; written at the end of round 1 without GPU time left to
// measure it -- see DESIGN.md section 8, item 1a): the band-pair transform as eight independent 512-point
// transforms, one per warp.  Only for N = 4096 and band slices of at most N/8 + 1 samples.
template <int LOG2N, bool WARP_BANDS = false>
__global__ void __launch_bounds__(D4C_BODY_THREADS, 2) d4c_body_kernel(BodyParams p) {
  extern __shared__ double2 smem_raw[];
  constexpr int N = 1 << LOG2N, NC = N / 2, bins = NC + 1;
  constexpr int binsp = (bins + 1) & ~1;  // keep 16-byte alignment
  cplx *S = smem_raw;
  double *SC = reinterpret_cast<double *>(S + wb_fft_slots(N));
  double *SP = SC + binsp;
  double *red = SP + binsp;                                       // 320
  unsigned long long *ctl = reinterpret_cast<unsigned long long *>(red + 320);   // select: control words + 2 x SEL_LIST keys
  double *coarse = reinterpret_cast<double *>(ctl + SEL_CTL_WORDS + 2 * SEL_LIST);  // D4C_MAX_AP + 2
  double *W = reinterpret_cast<double *>(S);
  double *seg = W;                                                // 2 * slots(N) doubles available
  const int seg_capacity = 2 * wb_fft_slots(N);
  double *win_hi = reinterpret_cast<double *>(S + wb_fft_slots(NC) + 8);  // >= N doubles above the packed real data

  const int frame = p.frame_begin + blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const double f0_in = p.f0[frame];
  if (f0_in == 0 || p.ap0[frame] <= p.threshold) {
    // not analysed: the row keeps its initial value 1 - kMySafeGuardMinimum (d4c.cpp:127-132, :146)
    const int out_bins = p.out_fft_size / 2 + 1;
    double *out = p.ap + (size_t)frame * out_bins;
    for (int i = tid; i < out_bins; i += nt) out[i] = 1.0 - WB_SAFEGUARD;
    return;
  }
  const double f0 = f0_in > WB_FLOOR_F0_D4C ? f0_in : WB_FLOOR_F0_D4C;
  const int fs = p.fs;
  const double pos = p.tpos[frame];
  const double *noise = p.noise + p.noise_off[frame];
  constexpr int log2n = LOG2N;

  // ---- static centroid: two windows at pos -/+ 0.25/f0 (d4c.cpp:339-405)
  // The reference transforms w[n] and (n+1) w[n] separately; both are real, so one complex
  // transform of z[n] = w[n] + i (n+1) w[n] yields both spectra.  While the window is being
  // built the imaginary parts of the slots hold the window samples.
  // (the window angle of sample j is the same in both windows: its rotation constants are set up once)
  double rot_sd = 0.0, rot_cd = 1.0, rot_sn = 0.0, rot_cs = 1.0;
  if constexpr (N == 16 * D4C_BODY_THREADS) {
    const double c1 = 2.0 / 4.0 / fs;
    const double c2 = WB_PI * f0;
    sincos(c2 * c1 * D4C_BODY_THREADS, &rot_sd, &rot_cd);
    sincos(c2 * (c1 * (tid - d4c_half_window(4.0, fs, f0))), &rot_sn, &rot_cs);
  }
  for (int c = 0; c < ((p.debug_skip & 8) ? 0 : 2); ++c) {
    const double cpos = (c == 0) ? pos - 0.25 / f0 : pos + 0.25 / f0;
    if constexpr (N == 16 * D4C_BODY_THREADS) {
      // One radix-16 butterfly per thread in the first FFT pass: thread t owns samples t + 256 q, exactly the
      // stride of the window loop, so the windowed, mean-free, unit-power waveform (d4c.cpp:246-303, :372-380)
      // is built in registers and handed to the transform without touching shared memory.
      const int hw = d4c_half_window(4.0, fs, f0);
      const int wlen = 2 * hw + 1;
      const int origin = wb_round(cpos * fs + 0.001);
      const double sd = rot_sd, cd = rot_cd;
      double sn = rot_sn, cs = rot_cs;
      double v[16], w[16];
      double s1 = 0.0, s2 = 0.0;
      // all waveform / noise loads first (independent: they overlap), parked in v[] / w[] until they are used
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int j = tid + q * D4C_BODY_THREADS;
        v[q] = 0.0; w[q] = 0.0;
        if (j < wlen) {
          v[q] = p.x[wb_min_i(p.x_length - 1, wb_max_i(0, origin + j - hw))];
          w[q] = noise[j];
        }
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int j = tid + q * D4C_BODY_THREADS;
        if (j < wlen) {
          const double wq = 0.42 + 0.5 * cs + 0.08 * (2.0 * cs * cs - 1.0);
          const double vq = v[q] * wq + w[q] * WB_SAFEGUARD;
          v[q] = vq; w[q] = wq;
          s1 += vq;
          s2 += wq;
          const double c_next = cs * cd - sn * sd;
          sn = sn * cd + cs * sd;
          cs = c_next;
        }
      }
      wb_block_sum2(s1, s2, red);
      const double coef = s1 / s2;
      double pw = 0.0;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        if (tid + q * D4C_BODY_THREADS < wlen) { v[q] -= w[q] * coef; pw += v[q] * v[q]; }
      }
      // unit power (d4c.cpp:372-380): one reciprocal per window instead of a division per sample (the
      // quotients differ from the reference's by at most one ulp)
      const double inv_power = 1.0 / sqrt(wb_block_sum(pw, red));
      noise += wlen;
      wb_cfft_dif_in_t<1, LOG2N>(S, p.tw_2n, [&](int j, int q) {
        cplx z = make_double2(0.0, 0.0);
        if (j < wlen) { z.x = v[q] * inv_power; z.y = z.x * (j + 1.0); }
        return z;
      });
    } else {
    const int wlen = d4c_windowed_waveform(p.x, p.x_length, fs, f0, cpos, D4C_BLACKMAN, 4.0, noise,
                                           [&](int j) -> double & { return S[wb_sidx(j)].y; }, red,
                                           [&](int j) -> double & { return S[wb_sidx(j)].x; });
    noise += wlen;
    double pw = 0.0;
    for (int j = tid; j < wlen; j += nt) { const double v = S[wb_sidx(j)].x; pw += v * v; }
    const double power = sqrt(wb_block_sum(pw, red));
    for (int j = tid; j < N; j += nt) {
      cplx z = make_double2(0.0, 0.0);
      if (j < wlen) { z.x = S[wb_sidx(j)].x / power; z.y = z.x * (j + 1.0); }
      S[wb_sidx(j)] = z;
    }
    __syncthreads();
    wb_cfft_dif_t<1, LOG2N, 16>(S, p.tw_2n);
    }
    for (int k = tid; k <= NC; k += nt) {
      const cplx zk = S[wb_sidx(wb_brev(k, log2n))];
      const cplx zc = S[wb_sidx(wb_brev((N - k) & (N - 1), log2n))];
      const double x1r = 0.5 * (zk.x + zc.x), x1i = 0.5 * (zk.y - zc.y);   // spectrum of w
      const double x2r = 0.5 * (zk.y + zc.y), x2i = -0.5 * (zk.x - zc.x);  // spectrum of (n+1) w
      const double cen = x2r * x1r + x1i * x2i;  // d4c.cpp:400
      SC[k] = (c == 0) ? cen : SC[k] + cen;
    }
    __syncthreads();
  }
  wb_dc_correction(SC, f0, fs, N);

  // ---- smoothed power spectrum (d4c.cpp:411-434)
  {
    const double rot[4] = {rot_sd, rot_cd, rot_sn, rot_cs};
    const int wlen = d4c_windowed_waveform(p.x, p.x_length, fs, f0, pos, D4C_HANNING, 4.0, noise,
                                           [&](int j) -> double & { return win_hi[j]; }, red,
                                           [&](int j) -> double & { return W[wb_didx(j)]; },
                                           (N == 16 * D4C_BODY_THREADS) ? rot : nullptr);
    for (int j = wlen + tid; j < N; j += nt) W[wb_didx(j)] = 0.0;
    __syncthreads();
    wb_rfft_t<1, LOG2N - 1, 16>(S, p.tw_n, [&](int k, cplx X) { SP[k] = X.x * X.x + X.y * X.y; });
    wb_dc_correction(SP, f0, fs, N);
    if (!wb_linear_smoothing(SP, SP, f0, fs, N, seg, seg_capacity, red)) {
      if (tid == 0) atomicExch(p.error_flag, WB_ERR_UNSUPPORTED);
      return;
    }
  }

  // ---- static group delay (d4c.cpp:440-460)
  for (int k = tid; k < bins; k += nt) SC[k] = SC[k] / SP[k];
  __syncthreads();
  if (!(p.debug_skip & 2)) {
    wb_linear_smoothing(SC, SC, f0 / 2.0, fs, N, seg, seg_capacity, red);
    wb_linear_smoothing(SC, SP, f0, fs, N, seg, seg_capacity, red);
  }
  for (int k = tid; k < bins; k += nt) SC[k] -= SP[k];
  __syncthreads();

  // ---- coarse aperiodicity per 3 kHz band (d4c.cpp:466-503)
  const int wl = p.window_length, hwl = wl / 2;
  const int boundary = wb_round(N * 8.0 / wl);
  const int m_small = bins - boundary - 1;  // cumulative-sum index bins - boundary - 2
  // Bands are transformed in pairs: band b in the real part, band b+1 in the imaginary part of ONE
  // complex N-point transform (both are real, so the spectra separate exactly); the two power
  // spectra are left interleaved in the FFT slots and both order statistics are resolved together.
  for (int b = 0; b < ((p.debug_skip & 4) ? 0 : p.n_ap); b += 2) {
    const bool has2 = (b + 1 < p.n_ap);
    const int center_a = static_cast<int>(WB_FREQ_INTERVAL * (b + 1) * N / fs);
    const int center_b = static_cast<int>(WB_FREQ_INTERVAL * (b + 2) * N / fs);
    // counters of the select's first digit (SP is free by now: 2 x 2^LOG2N packed 16-bit counters); the
    // barriers inside the transform order the clear before the counting
    {
      int4 *h4 = reinterpret_cast<int4 *>(SP);
      for (int i = tid; i < N / 4; i += nt) h4[i] = make_int4(0, 0, 0, 0);
    }
    if constexpr (WARP_BANDS && LOG2N == 12) {
      // The slice z[n] has at most N/8 + 1 = 513 non-zero samples, so Z[8 q + r] = sum_n (z[n] W_N^{n r}) W_512^{n q}
      // (+ z[512] W_N^{512 r}, which does not depend on q: it joins sample 0): warp r transforms the slice
      // modulated by W_N^{n r} with a 512-point FFT of its own -- lane l holds samples l + 32 t, a radix-16
      // butterfly over t, an exchange through a warp-private staging area, a radix-16 butterfly over the 32-point
      // sub-transforms' even / odd halves and a final radix-2 by shuffle.  No block barrier inside; ¾ of the
      // butterflies of the padded 4096-point transform.  Output in NATURAL order (model: profiles/band_fft_model.py).
      const int r = tid >> 5, l = tid & 31;
      const cplx *T = p.tw_2n;                       // e^{2 pi i k / 8192}
      cplx a[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const int n = l + 32 * t;
        cplx z = make_double2(0.0, 0.0);
        if (n < wl) {
          const double nw = __ldg(&p.nuttall[n]);
          z.x = SC[center_a - hwl + n] * nw;
          if (has2) z.y = SC[center_b - hwl + n] * nw;
        }
        if (r) z = wb_cmul(z, wb_tw<1>(T, (2 * n * r) & (2 * N - 1)));
        a[t] = z;
      }
      if (l == 0 && wl > 512) {                      // sample 512 folds onto sample 0 of every residue
        const double nw = __ldg(&p.nuttall[512]);
        cplx z2 = make_double2(SC[center_a - hwl + 512] * nw, has2 ? SC[center_b - hwl + 512] * nw : 0.0);
        z2 = wb_cmul(z2, wb_tw<1>(T, (1024 * r) & (2 * N - 1)));
        a[0] = wb_cadd(a[0], z2);
      }
      wb_dft16<1>(a);
      if (l) wb_apply_twiddles16<1>(T, 16 * l, a);   // a[p] *= W_512^{l p}
      cplx *stage = S + r * 544;                     // 16 sub-transforms x (32 + 2 padding) slots per warp
#pragma unroll
      for (int q = 0; q < 16; ++q) stage[q * 34 + l] = a[q];
      __syncwarp();
      const int sub = l >> 1, h = l & 1;             // lane pair (2 sub, 2 sub + 1) owns 32-point sub-transform `sub`
#pragma unroll
      for (int t = 0; t < 16; ++t) a[t] = stage[sub * 34 + h + 2 * t];
      wb_dft16<1>(a);
      if (h) wb_apply_twiddles16<1>(T, 256, a);      // a[p2] *= W_32^{p2}
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const double ox = __shfl_xor_sync(0xffffffffu, a[q].x, 1), oy = __shfl_xor_sync(0xffffffffu, a[q].y, 1);
        a[q] = h ? make_double2(ox - a[q].x, oy - a[q].y) : make_double2(a[q].x + ox, a[q].y + oy);
      }
      __syncthreads();                               // every warp has left its staging area
#pragma unroll
      for (int q = 0; q < 16; ++q) S[wb_sidx(8 * (sub + 16 * q + 256 * h) + r)] = a[q];
      __syncthreads();
    } else {
    // the windowed band slices feed the first FFT pass directly (the rest of the N points is zero padding)
    wb_cfft_dif_in_t<1, LOG2N>(S, p.tw_2n, [&](int j, int) {
      cplx z = make_double2(0.0, 0.0);
      if (j < wl) {
        const double nw = __ldg(&p.nuttall[j]);
        z.x = SC[center_a - hwl + j] * nw;
        if (has2) z.y = SC[center_b - hwl + j] * nw;
      }
      return z;
    });
    }
    // slot of spectrum bin k: natural order after the warp transforms, bit-reversed after the block-wide DIF
    auto bin_slot = [&](int k) { return (WARP_BANDS && LOG2N == 12) ? wb_sidx(k) : wb_sidx(wb_brev(k, log2n)); };
    double tot_a = 0.0, tot_b = 0.0;
    unsigned long long and_a = ~0ull, or_a = 0ull, and_b = ~0ull, or_b = 0ull;  // select keys of the power values
    unsigned *hist16 = reinterpret_cast<unsigned *>(SP);   // (SP is free from here on: counters of the select)
    for (int k = tid; k <= NC; k += nt) {
      const int slot = bin_slot(k);
      const cplx zk = S[slot];
      const cplx zc = S[bin_slot((N - k) & (N - 1))];
      const double ar = 0.5 * (zk.x + zc.x), ai = 0.5 * (zk.y - zc.y);   // spectrum of band b
      const double br = 0.5 * (zk.y + zc.y), bi = -0.5 * (zk.x - zc.x);  // spectrum of band b+1
      const double pa = ar * ar + ai * ai, pb = br * br + bi * bi;
      // slots of indices > NC are only read (by the thread that owns N - k), never written: in-place is safe
      S[slot] = make_double2(pa, pb);
      tot_a += pa;
      tot_b += pb;
      // first digit of the order-statistic select, counted while the values are at hand
      const unsigned long long ka = sel_key(pa), kb = sel_key(pb);
      and_a &= ka; or_a |= ka; and_b &= kb; or_b |= kb;
      sel_precount<LOG2N>(hist16, 0, ka);
      if (has2) sel_precount<LOG2N>(hist16, 1, kb);
    }
    wb_block_sum2(tot_a, tot_b, red);
    double low_a = 0.5 * tot_a, low_b = 0.5 * tot_b;
    if (!(p.debug_skip & 1))
    d4c_sum_smallest2<LOG2N - 1, LOG2N>([&](int i, int w) { const cplx v = S[bin_slot(i)]; return w == 0 ? v.x : v.y; },
                                        bins, m_small, has2, and_a, or_a, and_b, or_b, true, reinterpret_cast<int *>(SP), ctl, red,
                                        low_a, low_b);
    if (tid == 0) {
      const double rev = (f0 - 100) / 50.0;  // d4c.cpp:325-327
      double ca = 10 * log10(low_a / tot_a);
      ca = ca + rev;
      coarse[b + 1] = ca < 0.0 ? ca : 0.0;
      if (has2) {
        double cb = 10 * log10(low_b / tot_b);
        cb = cb + rev;
        coarse[b + 2] = cb < 0.0 ? cb : 0.0;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    coarse[0] = -60.0;                      // d4c.cpp:84
    coarse[p.n_ap + 1] = -WB_SAFEGUARD;     // d4c.cpp:85
  }
  __syncthreads();

  // ---- interp1 to the output bins + 10^(v/20) (d4c.cpp:160-168)
  const int out_bins = p.out_fft_size / 2 + 1;
  double *out = p.ap + (size_t)frame * out_bins;
  const int nk = p.n_ap + 2;
  for (int i = tid; i < out_bins; i += nt) {
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
               cudaStream_t stream, const WbRowChunks *chunks, const WbFrameRange *range, int phase, double *d_ap0_ext) {
  if (f0_length <= 0) return WB_OK;
  if (range && (range->begin < 0 || range->end > f0_length || range->begin > range->end || chunks)) return WB_ERR_ARG;
  if (phase < 0 || phase > 2 || (phase != 0 && !d_ap0_ext)) return WB_ERR_ARG;
  const int row0 = range ? range->begin : 0;
  const int n_rows = range ? range->end - range->begin : f0_length;
  const int N = wb_d4c_fft_size(fs), N_lt = wb_d4c_lt_fft_size(fs);
  const int l = ilog2_exact(N), l_lt = ilog2_exact(N_lt);
  if (l < 7 || N > 8192 || l_lt < 7 || N_lt > 16384) return WB_ERR_UNSUPPORTED;
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
  WB_LAUNCH("lt_count_scan_kernel", lt_count_scan_kernel<<<1, 1024, 0, stream>>>(d_f0, f0_length, fs, 40.0, d_offsets, nullptr, d_lt_total));
  WB_CUDA_CHECK(cudaGetLastError());
  if (rng.wait_skip_in) WB_CUDA_CHECK(cudaStreamWaitEvent(stream, rng.wait_skip_in, 0));
  int rc;
  if (phase != 2) {
  if (range) {
    if ((rc = wb_range_offsets(d_offsets, *range, d_rel, rng.skip_in, d_pos, d_pos + 1, stream))) return rc;
    if (n_rows > 0 && (rc = wb_rng_fill(rng.state, d_pos, d_pos + 1, max_noise_lt, d_noise, stream))) return rc;
  } else if ((rc = wb_rng_fill(rng.state, rng.skip_in, d_offsets + f0_length, max_noise_lt, d_noise, stream))) return rc;
  if (n_rows > 0) {
    LtParams p;
    p.x = d_x; p.x_length = x_length; p.tpos = d_tpos; p.f0 = d_f0; p.f0_length = f0_length;
    p.fs = fs; p.fft_size = N_lt; p.log2nc = l_lt - 1; p.lowest_f0 = 40.0;
    p.boundary0 = static_cast<int>(ceil(100.0 * N_lt / fs));
    p.boundary1 = static_cast<int>(ceil(4000.0 * N_lt / fs));
    p.boundary2 = static_cast<int>(ceil(7900.0 * N_lt / fs));
    p.twiddle = tw_lt; p.noise = d_noise; p.noise_off = range ? d_rel : d_offsets; p.ap0 = d_ap0;
    p.frame_begin = row0;
    const size_t smem = sizeof(cplx) * wb_fft_slots(N_lt / 2) + sizeof(double) * (N_lt + 128);
    rc = WB_DISPATCH_LOG2(l_lt, 9, 14, {
      if (cudaFuncSetAttribute(lt_frame_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WB_ERR_CUDA;
      WB_LAUNCH("lt_frame_kernel", lt_frame_kernel<L2><<<n_rows, 256, smem, stream>>>(p));
    });
    if (rc) return rc;
    WB_CUDA_CHECK(cudaGetLastError());
  }
  }
  if (phase == 1) return WB_OK;   // (the stream bookkeeping is done by the body phase)
  // ---- body
  WB_LAUNCH("body_count_scan_kernel", body_count_scan_kernel<<<1, 1024, 0, stream>>>(d_f0, d_ap0, f0_length, fs, threshold, d_offsets,
                                                                                  rng.skip_in, d_lt_total, d_skip_mid, d_skip_end));
  WB_CUDA_CHECK(cudaGetLastError());
  if (rng.record_skip_out) WB_CUDA_CHECK(cudaEventRecord(rng.record_skip_out, stream));
  if (range) {
    if ((rc = wb_range_offsets(d_offsets, *range, d_rel, d_skip_mid, d_pos, d_pos + 1, stream))) return rc;
    if (n_rows > 0 && (rc = wb_rng_fill(rng.state, d_pos, d_pos + 1, max_noise_body, d_noise, stream))) return rc;
  } else if ((rc = wb_rng_fill(rng.state, d_skip_mid, d_offsets + f0_length, max_noise_body, d_noise, stream))) return rc;
  if (n_rows > 0) {
    BodyParams p;
    p.x = d_x; p.x_length = x_length; p.tpos = d_tpos; p.f0 = d_f0; p.ap0 = d_ap0; p.f0_length = f0_length;
    p.fs = fs; p.fft_size_d4c = N; p.log2n = l; p.threshold = threshold;
    p.n_ap = n_ap; p.window_length = window_length; p.nuttall = d_nuttall;
    p.tw_n = tw_n; p.tw_2n = tw_2n; p.noise = d_noise; p.noise_off = range ? d_rel : d_offsets;
    p.out_fft_size = out_fft_size;
    p.ap = range ? d_ap - (size_t)row0 * (out_fft_size / 2 + 1) : d_ap;   // (rows are addressed by absolute frame)
    p.seg_capacity = 0;  // the smoothing scratch aliases the FFT slots
    p.error_flag = ws->error_flag();
    p.debug_skip = getenv("WB_D4C_SKIP") ? atoi(getenv("WB_D4C_SKIP")) : 0;
    const int binsp = ((N / 2 + 1) + 1) & ~1;
    const size_t smem = sizeof(cplx) * wb_fft_slots(N) + sizeof(double) * (2 * binsp + 320) +
                        sizeof(unsigned long long) * (SEL_CTL_WORDS + 2 * SEL_LIST) + sizeof(double) * (D4C_MAX_AP + 2);
    p.frame_begin = row0;
    static const bool warp_bands_env = getenv("WB_D4C_WARP_BANDS") && atoi(getenv("WB_D4C_WARP_BANDS")) != 0;
    if (warp_bands_env && l == 12 && window_length <= N / 8 + 1 && (!chunks || chunks->n <= 1)) {   // experimental, see the kernel
      if (cudaFuncSetAttribute(d4c_body_kernel<12, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WB_ERR_CUDA;
      WB_LAUNCH("d4c_body_kernel", d4c_body_kernel<12, true><<<n_rows, D4C_BODY_THREADS, smem, stream>>>(p));
      if (chunks && chunks->n == 1) WB_CUDA_CHECK(cudaEventRecord(chunks->ev[0], stream));
    } else if (!chunks || chunks->n <= 1) {
      rc = WB_DISPATCH_LOG2(l, 9, 13, {
        if (cudaFuncSetAttribute(d4c_body_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WB_ERR_CUDA;
        WB_LAUNCH("d4c_body_kernel", d4c_body_kernel<L2><<<n_rows, D4C_BODY_THREADS, smem, stream>>>(p));
      });
      if (rc) return rc;
      if (chunks && chunks->n == 1) WB_CUDA_CHECK(cudaEventRecord(chunks->ev[0], stream));
    } else {
      // row ranges on alternating streams (see WbRowChunks); everything before this point is on `stream`
      WB_CUDA_CHECK(cudaEventRecord(chunks->ev_ready, stream));
      WB_CUDA_CHECK(cudaStreamWaitEvent(chunks->alt, chunks->ev_ready, 0));
      for (int c = 0; c < chunks->n; ++c) {
        cudaStream_t cs = (c & 1) ? chunks->alt : stream;
        const int count = chunks->bounds[c + 1] - chunks->bounds[c];
        p.frame_begin = chunks->bounds[c];
        if (count > 0) {
          rc = WB_DISPATCH_LOG2(l, 9, 13, {
            if (cudaFuncSetAttribute(d4c_body_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WB_ERR_CUDA;
            WbLaunchScope scope("d4c_body_kernel", cs);
            d4c_body_kernel<L2><<<count, D4C_BODY_THREADS, smem, cs>>>(p);
          });
          if (rc) return rc;
        }
        WB_CUDA_CHECK(cudaEventRecord(chunks->ev[c], cs));
      }
      for (int c = 1; c < chunks->n; c += 2) WB_CUDA_CHECK(cudaStreamWaitEvent(stream, chunks->ev[c], 0));
    }
    WB_CUDA_CHECK(cudaGetLastError());
  }
  return rng.advance ? wb_rng_advance(rng.state, d_skip_end, nullptr, stream) : WB_OK;
}
