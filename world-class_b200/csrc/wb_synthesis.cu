// WORLD synthesis: time base -> pulse list -> one impulse response per pulse (7 FFTs in
// shared memory) -> deterministic overlap-add.
//
// Reference: /root/reference/src/synthesis.cpp
//   ctor :30-56, compute :77-177, getTimeBase :180-223, getTemporalParametersForTimeBase
//   :225-243, getPulseLocationsForTimeBase :245-288, getDCRemover :290-303,
//   getOneFrameSegment :308-344, getSpectralEnvelope :346-370, getAperiodicRatio :372-393,
//   getPeriodicResponse :403-437, getSpectrumWithFractionalTimeShift :443-457,
//   removeDCComponent :459-474, getAperiodicResponse :479-512, getNoiseSpectrum :514-530;
//   MinimumPhaseAnalysis::compute /root/reference/src/world_common.cpp:192-233.
//
// The reference's running phase sum (synthesis.cpp:257-264) is a sequential fp64
// recurrence whose rounding decides pulse sample indices.  phase_scan_kernel reproduces
// it BIT-EXACTLY in parallel: inside one binade every addition rounds to a multiple of a
// fixed ulp, so s <- fl(s + a) is an integer map S -> S + c(parity(S)) (the parity only
// matters for round-half-even ties); such maps compose associatively and are scanned.
// Binade crossings are detected and replayed with a genuine fp64 add.
#include "wb_internal.h"
#include "wb_fft.cuh"

#include <math.h>
#include <new>
#include <vector>

namespace {

// ---- K1: per-sample phase increment and VUV (synthesis.cpp:180-243) --------------------------
__global__ void timebase_kernel(const double *__restrict__ f0, int f0_length, int fs, double frame_period,
                                double lowest_f0, int y_length, double *__restrict__ incr,
                                unsigned char *__restrict__ vuv, int sample_begin, int out_offset) {
  // (streaming synthesis computes samples [sample_begin, y_length) into incr / vuv [out_offset ...])
  const int ii = sample_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (ii >= y_length) return;
  incr += out_offset - sample_begin;
  vuv += out_offset - sample_begin;
  const double t = ii / (double)fs;
  // histc over coarse_time_axis[k] = k * frame_period, k = 0..f0_length:
  // index = first k with coarse_time_axis[k] > t, clamped to [1, f0_length]
  int k = (int)(t / frame_period) + 1;
  if (k < 1) k = 1;
  if (k > f0_length) k = f0_length;
  while (k > 1 && !(t >= (k - 1) * frame_period)) --k;           // ensure x[k-1] <= t
  while (k < f0_length && !(t < k * frame_period)) ++k;           // ensure t < x[k] (or clamp)
  const double x0 = (k - 1) * frame_period, x1 = k * frame_period;
  const double s = (t - x0) / (x1 - x0);
  auto coarse_f0 = [&](int i) -> double {
    if (i < f0_length) { const double v = f0[i]; return (v < lowest_f0) ? 0.0 : v; }
    const double a = f0[f0_length - 1], b = f0[f0_length - 2];
    const double ca = (a < lowest_f0) ? 0.0 : a, cb = (b < lowest_f0) ? 0.0 : b;
    return ca * 2 - cb;
  };
  auto coarse_vuv = [&](int i) -> double {
    if (i < f0_length) { const double v = f0[i]; return (v < lowest_f0) ? 0.0 : 1.0; }
    const double a = f0[f0_length - 1], b = f0[f0_length - 2];
    const double va = (a < lowest_f0) ? 0.0 : 1.0, vb = (b < lowest_f0) ? 0.0 : 1.0;
    return va * 2 - vb;
  };
  const double f_lo = coarse_f0(k - 1), f_hi = coarse_f0(k);
  const double v_lo = coarse_vuv(k - 1), v_hi = coarse_vuv(k);
  const double fi = f_lo + s * (f_hi - f_lo);
  const double vi = v_lo + s * (v_hi - v_lo);
  const bool voiced = vi > 0.5;
  const double f = voiced ? fi : WB_DEFAULT_F0;
  const double two_pi = 2.0 * WB_PI;
  const double const_val = two_pi / fs;
  incr[ii] = f * const_val;
  vuv[ii] = voiced ? 1 : 0;
}

// ---- K2: exact parallel emulation of the sequential fp64 running sum -------------------------
#define PS_THREADS 1024
#define PS_PER 8
#define PS_CHUNK (PS_THREADS * PS_PER)

struct IncFn { unsigned long long ce, co; };  // increment if S is even / odd

__device__ __forceinline__ IncFn ps_compose(IncFn f, IncFn g) {  // apply f, then g
  IncFn h;
  h.ce = f.ce + ((f.ce & 1ull) ? g.co : g.ce);
  h.co = f.co + (((f.co + 1ull) & 1ull) ? g.co : g.ce);
  return h;
}

// increment function of adding `a` to a sum in binade e (ulp 2^(e-52)); flags elements that
// cannot be handled inside the binade model by returning a huge increment.
__device__ __forceinline__ IncFn ps_incfn(double a, int e) {
  const unsigned long long HUGE_INC = 1ull << 53;
  IncFn f;
  const long long bits = __double_as_longlong(a);
  const int ea = (int)((bits >> 52) & 0x7ff) - 1023;
  if (!(a > 0.0) || ea == -1023 || ea == 1024) { f.ce = f.co = HUGE_INC; return f; }  // <=0, subnormal, inf/nan
  const unsigned long long m = ((unsigned long long)bits & 0xfffffffffffffull) | (1ull << 52);
  const int shift = e - ea;
  if (shift < 0) { f.ce = f.co = HUGE_INC; return f; }
  if (shift == 0) { f.ce = f.co = m; return f; }
  if (shift > 62) { f.ce = f.co = 0ull; return f; }
  const unsigned long long q = m >> shift;
  const unsigned long long rem = m & ((1ull << shift) - 1ull);
  const unsigned long long half = 1ull << (shift - 1);
  if (rem > half) { f.ce = f.co = q + 1ull; }
  else if (rem < half) { f.ce = f.co = q; }
  else { f.ce = q + (q & 1ull); f.co = q + 1ull - (q & 1ull); }  // tie: round half to even
  return f;
}

// Block-wide EXCLUSIVE scan of per-thread composed increment functions (thread order = element order).
// `warp_fn` is a shared array of 32 entries; afterwards warp_fn[nwarps - 1] holds the composition of the whole
// block.  Contains two __syncthreads().
__device__ __forceinline__ IncFn ps_block_exclusive(IncFn acc, IncFn *warp_fn) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  IncFn incl = acc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    IncFn t;
    t.ce = __shfl_up_sync(0xffffffffu, incl.ce, o);
    t.co = __shfl_up_sync(0xffffffffu, incl.co, o);
    if (lane >= o) incl = ps_compose(t, incl);
  }
  if (lane == 31) warp_fn[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    IncFn wi = warp_fn[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      IncFn t;
      t.ce = __shfl_up_sync(0xffffffffu, wi.ce, o);
      t.co = __shfl_up_sync(0xffffffffu, wi.co, o);
      if (lane >= o) wi = ps_compose(t, wi);
    }
    warp_fn[lane] = wi;  // inclusive over warps
  }
  __syncthreads();
  IncFn pre; pre.ce = 0ull; pre.co = 0ull;
  if (warp > 0) pre = warp_fn[warp - 1];
  IncFn lane_excl;
  lane_excl.ce = __shfl_up_sync(0xffffffffu, incl.ce, 1);
  lane_excl.co = __shfl_up_sync(0xffffffffu, incl.co, 1);
  if (lane > 0) pre = ps_compose(pre, lane_excl);
  return pre;
}

// The scan runs in five launches so that only O(n / PS_CHUNK) work is sequential (one hour of audio is
// 1.7e8 samples: a single sequential pass was the Amdahl term of the sharded stream):
//   A  ps_chunk_sum_kernel    plain fp64 sum of every chunk of PS_CHUNK increments (parallel, approximate)
//   B  ps_binade_kernel       approximate running sum at the chunk boundaries -> the binade every chunk is
//                             EXPECTED to stay in (or "unknown" when a boundary is near)
//   C  ps_chunk_fn_kernel     the chunk's composed increment function for that binade (parallel, exact)
//   D  ps_sequential_kernel   the exact running sum chunk by chunk: one composed function per chunk where the
//                             expectation holds for the exact sum (verified), otherwise the chunk is walked
//                             cooperatively with genuine fp64 adds at binade crossings
//   E  ps_expand_kernel       the elements of the chunks D skipped over, from their exact start values (parallel)
// Chunk c covers elements [1 + c PS_CHUNK, 1 + (c + 1) PS_CHUNK); element 0 is the start value.
#define PS_E_UNKNOWN (-100000)

__global__ void __launch_bounds__(PS_THREADS) ps_chunk_sum_kernel(const double *__restrict__ incr, int n, double *__restrict__ chunk_sum) {
  __shared__ double red[32];
  const int begin = 1 + blockIdx.x * PS_CHUNK, end = min(n, begin + PS_CHUNK);
  double s = 0.0;
  for (int i = begin + threadIdx.x; i < end; i += PS_THREADS) s += incr[i];
  s = wb_block_sum(s, red);
  if (threadIdx.x == 0) chunk_sum[blockIdx.x] = s;
}

__device__ __forceinline__ int ps_exponent(double v) { return (int)((__double_as_longlong(v) >> 52) & 0x7ff) - 1023; }

__global__ void __launch_bounds__(PS_THREADS) ps_binade_kernel(const double *__restrict__ incr, const double *__restrict__ chunk_sum,
                                                               int n_chunks, int *__restrict__ e_pred) {
  __shared__ double s_warp[32];
  __shared__ double s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = incr[0];
  __syncthreads();
  for (int base = 0; base < n_chunks; base += PS_THREADS) {
    const int c = base + tid;
    const double a = c < n_chunks ? chunk_sum[c] : 0.0;
    double incl = a;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    double before = s_carry;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const double lo = before + (incl - a), hi = lo + a;   // approximate sums at the chunk's two ends
    if (c < n_chunks) {
      // relative error of these sums is far below 1e-9 (<= n eps); stay clear of binade boundaries by that much
      const double lo_m = lo * (1.0 - 1e-9), hi_m = hi * (1.0 + 1e-9);
      const int e_lo = ps_exponent(lo_m), e_hi = ps_exponent(hi_m);
      const bool ok = lo_m > 0.0 && a >= 0.0 && e_lo == e_hi && e_lo > -960 && e_lo < 1000;
      e_pred[c] = ok ? e_lo : PS_E_UNKNOWN;
    }
    __syncthreads();
    if (tid == PS_THREADS - 1) s_carry = hi;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(PS_THREADS) ps_chunk_fn_kernel(const double *__restrict__ incr, int n, const int *__restrict__ e_pred,
                                                                 IncFn *__restrict__ chunk_fn, int *__restrict__ chunk_bad) {
  __shared__ IncFn warp_fn[32];
  const int c = blockIdx.x;
  const int e = e_pred[c];
  if (e == PS_E_UNKNOWN) {
    if (threadIdx.x == 0) chunk_bad[c] = 1;
    return;
  }
  const int begin = 1 + c * PS_CHUNK, end = min(n, begin + PS_CHUNK);
  const int base = begin + threadIdx.x * PS_PER;
  IncFn acc; acc.ce = 0ull; acc.co = 0ull;
  int bad = 0;
#pragma unroll
  for (int q = 0; q < PS_PER; ++q) {
    const int i = base + q;
    if (i < end) {
      const IncFn f = ps_incfn(incr[i], e);
      bad |= (f.ce >= (1ull << 53)) ? 1 : 0;   // does not fit the binade model: the chunk is walked instead
      acc = ps_compose(acc, f);
    }
  }
  bad = __syncthreads_or(bad);
  (void)ps_block_exclusive(acc, warp_fn);
  if (threadIdx.x == 0) {
    chunk_fn[c] = warp_fn[PS_THREADS / 32 - 1];
    chunk_bad[c] = bad;
  }
}

// Writes total_phase[i] of the chunks it walks; the fmod / wrap detection is done by the (fully parallel)
// pulse kernels.
__global__ void __launch_bounds__(PS_THREADS) ps_sequential_kernel(const double *__restrict__ incr, int n,
                                                                   double *__restrict__ wrap_phase, int n_chunks,
                                                                   const int *__restrict__ e_pred, const IncFn *__restrict__ chunk_fn,
                                                                   const int *__restrict__ chunk_bad, double *__restrict__ chunk_start,
                                                                   int *__restrict__ chunk_expand) {
  __shared__ IncFn warp_fn[32];
  __shared__ int s_first_bad;
  __shared__ double s_sum;   // current running sum (exact double)
  __shared__ int s_pos;      // next element to process
  __shared__ int s_advanced;
  __shared__ int m_e[PS_THREADS];          // metadata of the next PS_THREADS chunks
  __shared__ IncFn m_fn[PS_THREADS];
  const int tid = threadIdx.x;
  if (tid == 0) {
    const double s0 = incr[0];  // total_phase[0] = interpolated_f0[0] * const_val
    wrap_phase[0] = s0;
    s_sum = s0;
    s_pos = 1;
  }
  __syncthreads();
  while (true) {
    const int pos = s_pos;
    if (pos >= n) break;
    const int c = (pos - 1) / PS_CHUNK;
    if ((pos - 1) % PS_CHUNK == 0) {
      // at a chunk boundary: skip over every chunk whose composed function applies to the exact sum
      {
        const int cc = c + tid;
        int e = PS_E_UNKNOWN;
        IncFn f; f.ce = 0ull; f.co = 0ull;
        if (cc < n_chunks && !chunk_bad[cc]) { e = e_pred[cc]; f = chunk_fn[cc]; }
        m_e[tid] = e;
        m_fn[tid] = f;
      }
      __syncthreads();
      if (tid == 0) {
        double sum = s_sum;
        int k = 0;
        const int k_max = min(PS_THREADS, n_chunks - c);
        while (k < k_max) {
          const int e = m_e[k];
          if (e == PS_E_UNKNOWN || !(sum > 0.0) || ps_exponent(sum) != e) break;
          const unsigned long long S = ((unsigned long long)__double_as_longlong(sum) & 0xfffffffffffffull) | (1ull << 52);
          const unsigned long long Sn = S + ((S & 1ull) ? m_fn[k].co : m_fn[k].ce);
          if (Sn >= (1ull << 53)) break;   // would leave the binade inside the chunk (increments are >= 0: monotone)
          chunk_start[c + k] = sum;
          chunk_expand[c + k] = 1;
          sum = (double)Sn * __longlong_as_double((long long)(e - 52 + 1023) << 52);
          ++k;
        }
        s_sum = sum;
        s_pos = min(n, 1 + (c + k) * PS_CHUNK);
        s_advanced = k;
        if (k == 0) chunk_expand[c] = 0;   // walked below
      }
      __syncthreads();
      if (s_advanced > 0) continue;
    }
    const int chunk_end = min(n, 1 + (c + 1) * PS_CHUNK);
    const double sum = s_sum;
    const long long sbits = __double_as_longlong(sum);
    const int e = (int)((sbits >> 52) & 0x7ff) - 1023;
    const bool state_ok = (sum > 0.0) && e != -1023 && e != 1024;
    if (!state_ok) {
      // degenerate running sum (zero/negative/subnormal/non-finite): genuine sequential add
      if (tid == 0) {
        const double s = sum + incr[pos];
        wrap_phase[pos] = s;
        s_sum = s;
        s_pos = pos + 1;
      }
      __syncthreads();
      continue;
    }
    const unsigned long long S0 = ((unsigned long long)sbits & 0xfffffffffffffull) | (1ull << 52);
    if (tid == 0) s_first_bad = 0x7fffffff;
    // local composition
    const int base = pos + tid * PS_PER;
    IncFn fn[PS_PER];
    IncFn acc; acc.ce = 0ull; acc.co = 0ull;
#pragma unroll
    for (int q = 0; q < PS_PER; ++q) {
      const int i = base + q;
      if (i < chunk_end) fn[q] = ps_incfn(incr[i], e);
      else { fn[q].ce = 0ull; fn[q].co = 0ull; }
      acc = ps_compose(acc, fn[q]);
    }
    const IncFn pre = ps_block_exclusive(acc, warp_fn);
    unsigned long long S = S0 + ((S0 & 1ull) ? pre.co : pre.ce);
    const unsigned long long LIMIT = 1ull << 53;
    const double ulp = __longlong_as_double((long long)(e - 52 + 1023) << 52);  // 2^(e-52), e-52 > -1023 here
    const bool ulp_ok = (e - 52) > -1022;
    int my_bad = 0x7fffffff;
    if (S >= LIMIT || !ulp_ok) my_bad = base;  // prefix already left the binade
#pragma unroll
    for (int q = 0; q < PS_PER; ++q) {
      const int i = base + q;
      if (i < chunk_end && my_bad == 0x7fffffff) {
        S += (S & 1ull) ? fn[q].co : fn[q].ce;
        if (S >= LIMIT) my_bad = i;
        else wrap_phase[i] = (double)S * ulp;
      }
    }
    if (my_bad != 0x7fffffff && my_bad < chunk_end) atomicMin(&s_first_bad, my_bad);
    __syncthreads();
    const int first_bad = s_first_bad;
    if (first_bad >= chunk_end) {
      // whole range valid: the thread that owns its last element publishes the sum
      const int last = chunk_end - 1;
      if (last >= base && last < base + PS_PER) { s_sum = (double)S * ulp; s_pos = chunk_end; }
    } else {
      // replay element first_bad with a genuine fp64 add on top of the exact sum before it
      const int owner = (first_bad - pos) / PS_PER;
      if (tid == owner) {
        unsigned long long Sb = S0 + ((S0 & 1ull) ? pre.co : pre.ce);
        for (int q = 0; q < PS_PER; ++q) {
          const int i = base + q;
          if (i >= first_bad) break;
          Sb += (Sb & 1ull) ? fn[q].co : fn[q].ce;
        }
        const double before = (double)Sb * ulp;
        const double s = before + incr[first_bad];
        wrap_phase[first_bad] = s;
        s_sum = s;
        s_pos = first_bad + 1;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(PS_THREADS) ps_expand_kernel(const double *__restrict__ incr, int n, double *__restrict__ wrap_phase,
                                                               const int *__restrict__ e_pred, const double *__restrict__ chunk_start,
                                                               const int *__restrict__ chunk_expand, int chunk_offset) {
  __shared__ IncFn warp_fn[32];
  const int c = blockIdx.x + chunk_offset;
  if (!chunk_expand[c]) return;
  const int e = e_pred[c];
  const int begin = 1 + c * PS_CHUNK, end = min(n, begin + PS_CHUNK);
  const int base = begin + threadIdx.x * PS_PER;
  const unsigned long long S0 = ((unsigned long long)__double_as_longlong(chunk_start[c]) & 0xfffffffffffffull) | (1ull << 52);
  IncFn fn[PS_PER];
  IncFn acc; acc.ce = 0ull; acc.co = 0ull;
#pragma unroll
  for (int q = 0; q < PS_PER; ++q) {
    const int i = base + q;
    if (i < end) fn[q] = ps_incfn(incr[i], e);
    else { fn[q].ce = 0ull; fn[q].co = 0ull; }
    acc = ps_compose(acc, fn[q]);
  }
  const IncFn pre = ps_block_exclusive(acc, warp_fn);
  unsigned long long S = S0 + ((S0 & 1ull) ? pre.co : pre.ce);
  const double ulp = __longlong_as_double((long long)(e - 52 + 1023) << 52);
#pragma unroll
  for (int q = 0; q < PS_PER; ++q) {
    const int i = base + q;
    if (i < end) {
      S += (S & 1ull) ? fn[q].co : fn[q].ce;
      wrap_phase[i] = (double)S * ulp;   // (D verified that the chunk's last sum stays inside the binade)
    }
  }
}

// ---- K3: pulse detection + ordered compaction (synthesis.cpp:266-281) -------------------------
#define PD_THREADS 256
__global__ void pulse_count_kernel(const double *__restrict__ total, int y_length,
                                   unsigned long long *__restrict__ block_counts) {
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const int ii = blockIdx.x * PD_THREADS + threadIdx.x;
  const double two_pi = 2.0 * WB_PI;
  bool flag = false;
  if (ii < y_length - 1) flag = fabs(fmod(total[ii + 1], two_pi) - fmod(total[ii], two_pi)) > WB_PI;
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt, __popc(m));
  __syncthreads();
  if (threadIdx.x == 0) block_counts[blockIdx.x] = (unsigned long long)s_cnt;
}

__global__ void pulse_write_kernel(const double *__restrict__ total, int y_length, int fs,
                                   const unsigned long long *__restrict__ block_offsets,
                                   int *__restrict__ pulse_index, double *__restrict__ pulse_shift, int max_pulses,
                                   int index_offset, int slot_offset, const unsigned char *__restrict__ vuv_local,
                                   unsigned char *__restrict__ pulse_vuv) {
  // (streaming synthesis appends: `total` is a piece of the stream whose element 0 is sample index_offset, the
  // pulses go to slots slot_offset + ..., and the voicing flag of the pulse's sample is kept per pulse)
  __shared__ int warp_cnt[PD_THREADS / 32];
  const int ii = blockIdx.x * PD_THREADS + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool flag = false;
  double w0 = 0.0, w1 = 0.0;
  if (ii < y_length - 1) {
    w0 = fmod(total[ii], 2.0 * WB_PI); w1 = fmod(total[ii + 1], 2.0 * WB_PI);
    flag = fabs(w1 - w0) > WB_PI;
  }
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) warp_cnt[warp] = __popc(m);
  __syncthreads();
  int before = 0;
  for (int w = 0; w < warp; ++w) before += warp_cnt[w];
  if (flag) {
    const long long slot = (long long)slot_offset + (long long)block_offsets[blockIdx.x] + before + __popc(m & ((1u << lane) - 1u));
    if (slot < max_pulses) {
      const double two_pi = 2.0 * WB_PI;
      const double y1 = w0 - two_pi;
      const double y2 = w1;
      const double x = -y1 / (y2 - y1);
      pulse_index[slot] = ii + index_offset;
      pulse_shift[slot] = x / fs;
      if (pulse_vuv) pulse_vuv[slot] = vuv_local[ii];
    }
  }
}

// total number of pulses and of randn() calls (sum of noise_size = idx[P-1] - idx[0])
__global__ void pulse_finalize_kernel(const unsigned long long *__restrict__ block_offsets, int n_blocks,
                                      const int *__restrict__ pulse_index, int max_pulses,
                                      int *__restrict__ n_pulses, unsigned long long *__restrict__ noise_count,
                                      int *__restrict__ error_flag) {
  unsigned long long P = block_offsets[n_blocks];
  if (P > (unsigned long long)max_pulses) { P = max_pulses; atomicExch(error_flag, WB_ERR_UNSUPPORTED); }
  n_pulses[0] = (int)P;
  n_pulses[1] = P >= 1 ? pulse_index[0] : 0;   // the sample the randn() stream of the excitation starts at
  *noise_count = (P >= 2) ? (unsigned long long)(pulse_index[P - 1] - pulse_index[0]) : 0ull;
}

// Range-restricted time base (one rank of a sharded stream): the pulse list holds the pulses of a sample window
// only, so the first and the last pulse of the WHOLE stream -- the origin of the excitation's randn() positions and
// the number of draws the stream makes (synthesis.cpp:520) -- are found in its first / last `span` samples.
__global__ void __launch_bounds__(1024) pulse_ends_kernel(const double *__restrict__ total, int y_length, int span,
                                                          const unsigned long long *__restrict__ block_offsets, int n_blocks,
                                                          int max_pulses, int *__restrict__ n_pulses,
                                                          unsigned long long *__restrict__ noise_count, int *__restrict__ error_flag) {
  __shared__ int s_first, s_last;
  if (threadIdx.x == 0) { s_first = 0x7fffffff; s_last = -1; }
  __syncthreads();
  const double two_pi = 2.0 * WB_PI;
  auto pulse_at = [&](int ii) { return fabs(fmod(total[ii + 1], two_pi) - fmod(total[ii], two_pi)) > WB_PI; };
  const int head_end = min(span, y_length - 1), tail_begin = max(0, y_length - 1 - span);
  for (int ii = threadIdx.x; ii < head_end; ii += blockDim.x) if (pulse_at(ii)) { atomicMin(&s_first, ii); break; }
  for (int ii = y_length - 2 - threadIdx.x; ii >= tail_begin; ii -= blockDim.x) if (pulse_at(ii)) { atomicMax(&s_last, ii); break; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long P = block_offsets[n_blocks];
    if (P > (unsigned long long)max_pulses) { P = max_pulses; atomicExch(error_flag, WB_ERR_UNSUPPORTED); }
    if (s_first == 0x7fffffff || s_last < 0) { atomicExch(error_flag, WB_ERR_UNSUPPORTED); s_first = 0; s_last = 0; }
    n_pulses[0] = (int)P;
    n_pulses[1] = s_first;
    *noise_count = s_last > s_first ? (unsigned long long)(s_last - s_first) : 0ull;
  }
}

// ---- K5: impulse response per pulse -----------------------------------------------------------
struct RespParams {
  const double *sp; const double *ap; int f0_length;
  int fs; int fft_size; int log2n; double frame_period;
  const int *pulse_index; const double *pulse_shift; const unsigned char *vuv;
  const int *n_pulses;
  const double *noise;           // randn stream starting at the first pulse
  const double *dc_remover;      // fft_size doubles (only the first half is used, Q9)
  const cplx *tw_n;              // fft_size entries
  const cplx *tw_2n;             // 2*fft_size entries
  double *response;              // [max_resp_pulses][fft_size]
  int max_resp_pulses;
  int *error_flag;
  // sharded streams (wb_synthesis_render_range): pulses [range[0], range[1]) only, noise[0] is the draw of
  // sample index range[2], sp / ap address frame `row_begin`; null = every pulse
  const int *range;
  int row_begin;
  const unsigned char *pulse_vuv;   // streaming synthesis: voicing flag per pulse instead of per sample (null = use vuv)
};

// pulses whose response overlaps samples [sample_begin, sample_end), and the share of the noise stream
// they read: range = {first pulse, one past the last, sample index of the first pulse}; *noise_skip /
// *noise_count position the randn() fill
__global__ void pulse_range_kernel(const int *__restrict__ pulse_index, const int *__restrict__ n_pulses,
                                   const int *__restrict__ first_pulse, int fft_size,
                                   int sample_begin, int sample_end, const unsigned long long *__restrict__ skip_in,
                                   int *__restrict__ range, unsigned long long *__restrict__ noise_skip,
                                   unsigned long long *__restrict__ noise_count) {
  const int P = *n_pulses, half = fft_size / 2;
  // a pulse at idx covers samples (idx - half, idx - half + fft_size]: wanted are idx > sample_begin + half - 1 - fft_size
  // and idx <= sample_end - 1 + half - 1 (see ola_kernel)
  auto first_above = [&](int v) {   // first pulse with idx > v
    int lo = 0, hi = P;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (pulse_index[mid] > v) hi = mid; else lo = mid + 1;
    }
    return lo;
  };
  const int p_lo = first_above(sample_begin + half - 1 - fft_size);
  const int p_hi = first_above(sample_end - 1 + half - 1);
  range[0] = p_lo;
  range[1] = p_hi;
  if (P == 0 || p_lo >= p_hi) { range[2] = 0; *noise_skip = skip_in ? *skip_in : 0ull; *noise_count = 0ull; return; }
  const int idx_lo = pulse_index[p_lo];
  const int idx_end = pulse_index[p_hi < P ? p_hi : P - 1];   // the last pulse of the stream draws nothing (Q11)
  range[2] = idx_lo;
  *noise_skip = (skip_in ? *skip_in : 0ull) + (unsigned long long)(idx_lo - *first_pulse);   // (first pulse of the WHOLE stream)
  *noise_count = (unsigned long long)(idx_end - idx_lo);
}

// MinimumPhaseAnalysis::compute (world_common.cpp:192-233).  On entry the packed real view W of
// S holds log_spectrum[0..NC] (this function mirrors it); on exit MP[k], k = 0..NC, holds the
// minimum phase spectrum.  S needs wb_fft_slots(N / 2) slots (packed real transforms only).
template <int LOG2N, bool WL = false>
__device__ __forceinline__ void minimum_phase(cplx *S, cplx *MP, const cplx *tw_n, const cplx *tw_2n) {
  constexpr int N = 1 << LOG2N, NC = N / 2, log2n = LOG2N;
  double *W = reinterpret_cast<double *>(S);
  for (int i = NC + 1 + threadIdx.x; i < N; i += blockDim.x) W[wb_didx(i)] = W[wb_didx(N - i)];
  __syncthreads();
  // "inverse_fft" is a forward r2c in the reference; sign flips / doubling at :203-209
  // The mirrored log spectrum is real and even, so its transform -- the cepstrum -- is real: the reference carries
  // imaginary parts that are pure rounding noise (~1e-16 of the real parts) into its complex N-point transform.
  // Dropping them makes the folded cepstrum a REAL sequence with a zero upper half, and its first N/2 + 1 bins
  // come from a real transform (a half-size complex FFT) instead of a full complex one.
  wb_rfft_t<1, LOG2N - 1, WB_FFT_DEFAULT_RADIX, WL>(S, tw_n, [&](int k, cplx X) {
    MP[k].x = (k == 0 || k == NC) ? X.x : X.x * 2.0;
  });
  for (int i = threadIdx.x; i < N; i += blockDim.x) W[wb_didx(i)] = (i <= NC) ? MP[i].x : 0.0;
  __syncthreads();
  wb_rfft_t<1, LOG2N - 1, WB_FFT_DEFAULT_RADIX, WL>(S, tw_n, [&](int k, cplx X) {   // first half of the reference's c2c FFT_FORWARD
    const double tmp = exp(X.x / N);
    double sn, cs;
    sincos(X.y / N, &sn, &cs);
    MP[k] = make_double2(tmp * cs, tmp * sn);
  });
  __syncthreads();
  (void)tw_2n; (void)log2n;
}

template <int LOG2N>
__global__ void __launch_bounds__(256, 4) response_kernel(RespParams p) {
  extern __shared__ double2 smem_raw[];
  constexpr int N = 1 << LOG2N, NC = N / 2, bins = NC + 1;
  constexpr bool WL = (NC / 8 <= 256);        // one radix-8 butterfly per thread and pass (256 threads): warp-local late passes
  const int binsp = (bins + 1) & ~1;
  cplx *S = smem_raw;                         // wb_fft_slots(NC): every transform here is a real one (packed N/2-point complex)
  cplx *MP = S + wb_fft_slots(NC);            // bins: minimum-phase spectrum (x noise spectrum for the aperiodic part)
  double *SE = reinterpret_cast<double *>(MP + binsp);  // spectral envelope
  double *AR = SE + binsp;                    // aperiodic ratio
  double *red = AR + binsp;                   // 128
  double *W = reinterpret_cast<double *>(S);
  const int tid = threadIdx.x, nt = blockDim.x;
  const int P = *p.n_pulses;
  const int p_lo = p.range ? p.range[0] : 0, p_hi = p.range ? p.range[1] : P;
  if (p_hi - p_lo > p.max_resp_pulses) {  // f0 exceeded the caller's bound: refuse rather than overrun
    if (tid == 0 && blockIdx.x == 0) atomicExch(p.error_flag, WB_ERR_ARG);
    return;
  }

  for (int pulse = p_lo + blockIdx.x; pulse < p_hi; pulse += gridDim.x) {
    double *resp = p.response + (size_t)(pulse - p_lo) * N;
    const int idx = p.pulse_index[pulse];
    const int idx_next = p.pulse_index[wb_min_i(P - 1, pulse + 1)];
    const int noise_size = idx_next - idx;
    if (noise_size <= 0) {  // last pulse: 0 * periodic + 0 (synthesis.cpp:339-343)
      for (int i = tid; i < N; i += nt) resp[i] = 0.0;
      continue;
    }
    const double current_vuv = (p.pulse_vuv ? p.pulse_vuv[pulse] : p.vuv[idx]) ? 1.0 : 0.0;
    const double current_time = idx / (double)p.fs;
    const double frac_shift = p.pulse_shift[pulse];

    // ---- interpolated spectral envelope / aperiodic ratio (synthesis.cpp:346-398)
    const double tf = current_time / p.frame_period;
    const int fl = wb_min_i(p.f0_length - 1, (int)floor(tf));
    const int cl = wb_min_i(p.f0_length - 1, (int)ceil(tf));
    const double interp = tf - fl;
    const double *sp_f = p.sp + (size_t)(fl - p.row_begin) * bins, *sp_c = p.sp + (size_t)(cl - p.row_begin) * bins;
    const double *ap_f = p.ap + (size_t)(fl - p.row_begin) * bins, *ap_c = p.ap + (size_t)(cl - p.row_begin) * bins;
    for (int k = tid; k < bins; k += nt) {
      double se, ar;
      const double af = fmax(0.001, fmin(0.999999999999, ap_f[k]));
      if (fl == cl) {
        se = fabs(sp_f[k]);
        ar = af * af;
      } else {
        const double ac = fmax(0.001, fmin(0.999999999999, ap_c[k]));
        se = (1.0 - interp) * fabs(sp_f[k]) + interp * fabs(sp_c[k]);
        const double a = (1.0 - interp) * af + interp * ac;
        ar = a * a;
      }
      SE[k] = se;
      AR[k] = ar;
    }
    __syncthreads();

    // ---- periodic response (synthesis.cpp:403-474)
    const bool periodic_on = !(current_vuv <= 0.5 || AR[0] > 0.999);
    if (periodic_on) {
      for (int k = tid; k < bins; k += nt) W[wb_didx(k)] = log(SE[k] * (1.0 - AR[k]) + WB_SAFEGUARD) / 2.0;
      __syncthreads();
      minimum_phase<LOG2N, WL>(S, MP, p.tw_n, p.tw_2n);
      const double coefficient = 2.0 * WB_PI * frac_shift * p.fs / N;
      wb_irfft_t<-1, LOG2N - 1, WB_FFT_DEFAULT_RADIX, WL>(S, p.tw_n, [&](int k) {
        const cplx v = MP[k];
        const double re2 = cos(coefficient * k);
        const double im2 = sqrt(1.0 - re2 * re2);  // Q8: always >= 0
        return make_double2(v.x * re2 - v.y * im2, v.x * im2 + v.y * re2);
      });
      // fftshift + removeDCComponent (Q9): first half <- -dc * r[i]; second half <- out[i] - dc * r[i]
      double part = 0.0;
      for (int i = tid; i < NC; i += nt) part += W[wb_didx(i)];
      const double dc = wb_block_sum(part, red);
      // (parked in the pulse's global row; the same thread finishes element i below)
      for (int i = tid; i < NC; i += nt) {
        const double rm = -dc * p.dc_remover[i];
        resp[i] = rm;
        resp[i + NC] = W[wb_didx(i)] + rm;
      }
    } else {
      for (int i = tid; i < N; i += nt) resp[i] = 0.0;
    }
    __syncthreads();

    // ---- aperiodic response (synthesis.cpp:479-530)
    {
      // minimum phase of the aperiodic envelope first, then the noise spectrum is multiplied into it bin by
      // bin as the transform emits it (no separate noise-spectrum buffer)
      const double *nz = p.noise + (idx - (p.range ? p.range[2] : p.pulse_index[0]));
      if (current_vuv != 0.0) {
        for (int k = tid; k < bins; k += nt) W[wb_didx(k)] = log(SE[k] * AR[k]) / 2.0;
      } else {
        for (int k = tid; k < bins; k += nt) W[wb_didx(k)] = log(SE[k]) / 2.0;
      }
      __syncthreads();
      minimum_phase<LOG2N, WL>(S, MP, p.tw_n, p.tw_2n);
      double part = 0.0;
      for (int i = tid; i < noise_size; i += nt) part += nz[i];
      const double average = wb_block_sum(part, red) / noise_size;
      for (int i = tid; i < N; i += nt) W[wb_didx(i)] = (i < noise_size) ? nz[i] - average : 0.0;
      __syncthreads();
      wb_rfft_t<1, LOG2N - 1, WB_FFT_DEFAULT_RADIX, WL>(S, p.tw_n, [&](int k, cplx X) {
        const cplx a = MP[k];
        MP[k] = make_double2(a.x * X.x - a.y * X.y, a.x * X.y + a.y * X.x);
      });
      wb_irfft_t<-1, LOG2N - 1, WB_FFT_DEFAULT_RADIX, WL>(S, p.tw_n, [&](int k) { return MP[k]; });
    }
    // ---- combine (synthesis.cpp:339-343) with the aperiodic fftshift folded in
    const double sqrt_noise_size = sqrt((double)noise_size);
    for (int i = tid; i < N; i += nt) {
      const double aper = W[wb_didx(i < NC ? i + NC : i - NC)];
      resp[i] = (resp[i] * sqrt_noise_size + aper) / N;
    }
    __syncthreads();
  }
}

// ---- K6: deterministic overlap-add (synthesis.cpp:118-139) -------------------------------------
// out[n] = sum over pulses (in pulse order, like the serial reference) of
// response[p][n - (idx_p - N/2 + 1)].
// Sharded streams: samples [sample_begin, sample_end) only, out[0] is sample_begin and response[0] belongs to
// pulse range[0] (null = the whole waveform).
__global__ void ola_kernel(const double *__restrict__ response, const int *__restrict__ pulse_index,
                           const int *__restrict__ n_pulses, int max_resp_pulses, int fft_size, int out_length,
                           double *__restrict__ out, int sample_begin, int sample_end, const int *__restrict__ range) {
  const int n = sample_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= sample_end) return;
  out += -sample_begin;
  const int P = *n_pulses;
  const int p_first = range ? range[0] : 0;
  if ((range ? range[1] - range[0] : P) > max_resp_pulses) { out[n] = 0.0; return; }
  const int half = fft_size / 2;
  // pulses with 0 <= n - (idx - half + 1) < fft_size  <=>  n - half - ... :
  // idx in [n + half + 1 - fft_size, n + half - 1 + 1 - 0] -> idx >= n - half + 1 ... derive:
  // j = n - idx + half - 1 in [0, fft_size)  <=>  idx in (n + half - 1 - fft_size, n + half - 1]
  const int lo_val = n + half - 1 - fft_size;  // exclusive
  const int hi_val = n + half - 1;             // inclusive
  // first pulse with idx > lo_val
  int lo = 0, hi = P;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (pulse_index[mid] > lo_val) hi = mid; else lo = mid + 1;
  }
  double acc = 0.0;
  for (int q = lo; q < P; ++q) {
    const int idx = pulse_index[q];
    if (idx > hi_val) break;
    const int index = idx - half;
    // the reference skips pulses with index + fft_size < 0 or index + 1 >= out_length entirely
    if (index + fft_size < 0 || index + 1 >= out_length) continue;
    const int j = n - index - 1;
    acc += response[(size_t)(q - p_first) * fft_size + j];
  }
  out[n] = acc;
}

}  // namespace

// Host-side tables: dc_remover (synthesis.cpp:290-303), computed with host libm and the
// reference's sequential accumulate.
static void make_dc_remover(int fft_size, std::vector<double> &r) {
  r.assign(fft_size, 0.0);
  const double const_val = 2.0 * WB_PI / (1.0 + fft_size);
  for (int ii = 0; ii < fft_size / 2; ii++) r[ii] = 0.5 - 0.5 * cos(const_val * (ii + 1.0));
  double acc = 0.0;
  for (int ii = 0; ii < fft_size / 2; ii++) acc += r[ii];
  const double dc_component = acc * 2;
  for (int ii = 0; ii < fft_size / 2; ii++) {
    r[ii] /= dc_component;
    r[fft_size - ii - 1] = r[ii];
  }
}

// exact running sum total[i] = fl(total[i-1] + incr[i]), total[0] = incr[0] (see the ps_* kernels)
// `windows` (optional, up to 3 sample windows [lo, hi)): only the chunks that hold these samples are expanded -- the
// rest of d_total stays unwritten (range-restricted time base of a sharded stream).
static int run_phase_scan(WbWorkspace *ws, const double *d_incr, int out_length, double *d_total, cudaStream_t stream,
                          const int (*windows)[2] = nullptr, int n_windows = 0) {
  {
    const int n_chunks = out_length > 1 ? (out_length - 1 + PS_CHUNK - 1) / PS_CHUNK : 0;
    const int nc = n_chunks > 0 ? n_chunks : 1;
    double *d_csum = (double *)ws->get("syn_ps_sum", sizeof(double) * nc);
    double *d_cstart = (double *)ws->get("syn_ps_start", sizeof(double) * nc);
    int *d_ce = (int *)ws->get("syn_ps_e", sizeof(int) * nc);
    int *d_cbad = (int *)ws->get("syn_ps_bad", sizeof(int) * nc);
    int *d_cexp = (int *)ws->get("syn_ps_expand", sizeof(int) * nc);
    IncFn *d_cfn = (IncFn *)ws->get("syn_ps_fn", sizeof(IncFn) * nc);
    if (!d_csum || !d_cstart || !d_ce || !d_cbad || !d_cexp || !d_cfn) return WB_ERR_CUDA;
    if (n_chunks > 0) {
      WB_LAUNCH("ps_chunk_sum_kernel", ps_chunk_sum_kernel<<<n_chunks, PS_THREADS, 0, stream>>>(d_incr, out_length, d_csum));
      WB_LAUNCH("ps_binade_kernel", ps_binade_kernel<<<1, PS_THREADS, 0, stream>>>(d_incr, d_csum, n_chunks, d_ce));
      WB_LAUNCH("ps_chunk_fn_kernel", ps_chunk_fn_kernel<<<n_chunks, PS_THREADS, 0, stream>>>(d_incr, out_length, d_ce, d_cfn, d_cbad));
    }
    WB_LAUNCH("phase_scan_kernel", ps_sequential_kernel<<<1, PS_THREADS, 0, stream>>>(d_incr, out_length, d_total, n_chunks, d_ce, d_cfn,
                                                                                   d_cbad, d_cstart, d_cexp));
    if (n_chunks > 0 && !windows)
      WB_LAUNCH("ps_expand_kernel", ps_expand_kernel<<<n_chunks, PS_THREADS, 0, stream>>>(d_incr, out_length, d_total, d_ce, d_cstart, d_cexp, 0));
    for (int w = 0; n_chunks > 0 && w < n_windows; ++w) {
      // chunk c holds elements [1 + c PS_CHUNK, 1 + (c + 1) PS_CHUNK); element 0 is written by the sequential kernel
      const int lo = windows[w][0] < 1 ? 1 : windows[w][0], hi = windows[w][1] > out_length ? out_length : windows[w][1];
      if (hi <= lo) continue;
      const int c0 = (lo - 1) / PS_CHUNK, c1 = (hi - 2) / PS_CHUNK;
      WB_LAUNCH("ps_expand_kernel", ps_expand_kernel<<<c1 - c0 + 1, PS_THREADS, 0, stream>>>(d_incr, out_length, d_total, d_ce, d_cstart, d_cexp, c0));
    }
  }
  WB_CUDA_CHECK(cudaGetLastError());
  return WB_OK;
}

// ---- host side -------------------------------------------------------------------------------
// Part 1 (depends on f0 only): time base, exact phase scan, pulse list.  May run on a side
// stream while CheapTrick / D4C are still busy.
// sample_begin / sample_end (a rank of a sharded stream): the pulse list is built for the pulses that reach into
// [sample_begin, sample_end) only -- the per-sample passes over the rest of the stream (chunk expansion, pulse
// detection, compaction: two thirds of the time base) are skipped; sample_end < 0 = the whole stream.
int wb_synthesis_timebase(WbWorkspace *ws, int fs, int fft_size, double frame_period_ms, const double *d_f0,
                          int f0_length, int out_length, cudaStream_t stream, const WbRngCursor *noise_cursor,
                          int sample_begin, int sample_end) {
  if (out_length <= 0) return WB_OK;
  if (f0_length < 2) return WB_ERR_ARG;
  // pulses wanted: idx in (sample_begin + fft/2 - 1 - fft, sample_end + fft/2 - 2], plus the pulse after the last one
  // (its noise ends there); pulses are less than 2 fft_size samples apart (see render_range_core)
  const int span = 3 * fft_size + 64;   // (head / tail searched for the stream's first / last pulse: a pulse period is < 2 fft_size samples)
  int win_lo = 0, win_hi = out_length;
  if (sample_end >= 0) {
    if (sample_begin < 0 || sample_end > out_length || sample_begin > sample_end || noise_cursor) return WB_ERR_ARG;
    win_lo = sample_begin - fft_size - 8 < 0 ? 0 : sample_begin - fft_size - 8;
    win_hi = sample_end + 3 * fft_size + 8 > out_length || sample_end + 3 * fft_size + 8 < 0 ? out_length : sample_end + 3 * fft_size + 8;
  }
  const bool ranged = sample_end >= 0 && out_length > 4 * span && (win_lo > 0 || win_hi < out_length);
  if (!ranged) { win_lo = 0; win_hi = out_length; }
  const double frame_period = frame_period_ms / 1000.;      // synthesis.cpp:31
  const double lowest_f0 = fs / fft_size + 1.0;             // synthesis.cpp:97 (integer division)
  // Every pulse needs a 2 pi phase advance and one sample adds at most 2 pi max(f0, 500) / fs,
  // so pulses <= out_length * max(f0_max, 500) / fs + 1.  The index/shift lists are sized for
  // f0 <= fs / 4; the response buffer for the tighter bound (or the exact count).
  const int win = win_hi - win_lo;
  const int max_pulses = win / 4 + 16;
  double *d_incr = (double *)ws->get("syn_incr", sizeof(double) * out_length);
  double *d_total = (double *)ws->get("syn_wrap", sizeof(double) * out_length);
  unsigned char *d_vuv = (unsigned char *)ws->get("syn_vuv", out_length);
  const int n_blocks = (win + PD_THREADS - 1) / PD_THREADS;
  unsigned long long *d_bc = (unsigned long long *)ws->get("syn_bcount", sizeof(unsigned long long) * (n_blocks + 1));
  unsigned long long *d_bo = (unsigned long long *)ws->get("syn_boff", sizeof(unsigned long long) * (n_blocks + 1));
  int *d_pidx = (int *)ws->get("syn_pidx", sizeof(int) * max_pulses);
  double *d_pshift = (double *)ws->get("syn_pshift", sizeof(double) * max_pulses);
  int *d_np = (int *)ws->get("syn_np", sizeof(int) * 4);
  unsigned long long *d_ncount = (unsigned long long *)ws->get("syn_ncount", sizeof(unsigned long long));
  if (!d_incr || !d_total || !d_vuv || !d_bc || !d_bo || !d_pidx || !d_pshift || !d_np || !d_ncount) return WB_ERR_CUDA;
  WB_LAUNCH("timebase_kernel", timebase_kernel<<<(out_length + 255) / 256, 256, 0, stream>>>(
      d_f0, f0_length, fs, frame_period, lowest_f0, out_length, d_incr, d_vuv, 0, 0));
  {
    const int windows[3][2] = {{win_lo, win_hi}, {0, span + 1}, {out_length - span - 2, out_length}};
    const int rc_scan = ranged ? run_phase_scan(ws, d_incr, out_length, d_total, stream, windows, 3)
                               : run_phase_scan(ws, d_incr, out_length, d_total, stream);
    if (rc_scan) return rc_scan;
  }
  // (a window is a piece of the stream whose element 0 is sample win_lo, like a piece of a streaming synthesis)
  WB_LAUNCH("pulse_count_kernel", pulse_count_kernel<<<n_blocks, PD_THREADS, 0, stream>>>(d_total + win_lo, win, d_bc));
  int rc = wb_exclusive_scan_u64(d_bc, d_bo, n_blocks, stream, nullptr, nullptr, ws);
  if (rc) return rc;
  WB_LAUNCH("pulse_write_kernel", pulse_write_kernel<<<n_blocks, PD_THREADS, 0, stream>>>(d_total + win_lo, win, fs, d_bo, d_pidx, d_pshift, max_pulses,
                                                                                        win_lo, 0, nullptr, nullptr));
  if (ranged)
    WB_LAUNCH("pulse_ends_kernel", pulse_ends_kernel<<<1, 1024, 0, stream>>>(d_total, out_length, span, d_bo, n_blocks, max_pulses, d_np, d_ncount,
                                                                          ws->error_flag()));
  else
    WB_LAUNCH("pulse_finalize_kernel", pulse_finalize_kernel<<<1, 1, 0, stream>>>(d_bo, n_blocks, d_pidx, max_pulses, d_np, d_ncount, ws->error_flag()));
  WB_CUDA_CHECK(cudaGetLastError());
  if (noise_cursor) {
    // the aperiodic excitation only needs the pulse span: draw it here, off the critical path
    double *d_noise = (double *)ws->get("noise_syn", sizeof(double) * out_length);
    if (!d_noise) return WB_ERR_CUDA;
    if (noise_cursor->wait_skip_in) WB_CUDA_CHECK(cudaStreamWaitEvent(stream, noise_cursor->wait_skip_in, 0));
    return wb_rng_fill(noise_cursor->state, noise_cursor->skip_in, d_ncount, (unsigned long long)out_length, d_noise, stream);
  }
  return WB_OK;
}

// Part 2: noise, impulse responses, overlap-add.  Must follow wb_synthesis_timebase on `ws`.
// f0_upper_bound: an upper bound of max(f0) known to the host (e.g. Harvest's f0_ceil); <= 0 if
// unknown, in which case the pulse count is read back (one stream synchronisation).
int wb_synthesis_render(WbWorkspace *ws, int fs, int fft_size, double frame_period_ms, int f0_length,
                        const double *d_sp, const double *d_ap, int out_length, double *d_out,
                        double f0_upper_bound, const WbRngCursor &rng, cudaStream_t stream, bool noise_ready) {
  if (out_length <= 0) return WB_OK;
  int log2n = 0;
  while ((1 << log2n) < fft_size) ++log2n;
  if ((1 << log2n) != fft_size || fft_size < 128 || fft_size > 8192) return WB_ERR_UNSUPPORTED;
  const double frame_period = frame_period_ms / 1000.;
  const int max_pulses = out_length / 4 + 16;
  unsigned char *d_vuv = (unsigned char *)ws->find("syn_vuv");
  int *d_pidx = (int *)ws->find("syn_pidx");
  double *d_pshift = (double *)ws->find("syn_pshift");
  int *d_np = (int *)ws->find("syn_np");
  unsigned long long *d_ncount = (unsigned long long *)ws->find("syn_ncount");
  double *d_noise = (double *)ws->get("noise_syn", sizeof(double) * out_length);
  double *d_dcr = (double *)ws->get("syn_dcr", sizeof(double) * fft_size);
  if (!d_vuv || !d_pidx || !d_pshift || !d_np || !d_ncount || !d_noise || !d_dcr) return WB_ERR_CUDA;
  const cplx *tw_n = wb_twiddle_table(fft_size);
  const cplx *tw_2n = wb_twiddle_table(2 * fft_size);
  if (!tw_n || !tw_2n) return WB_ERR_CUDA;
  {
    // cached per workspace (see the Nuttall table in wb_d4c.cu)
    int *tag = (int *)ws->get_pinned("syn_dcr_tag", sizeof(int) * 4);
    if (!tag) return WB_ERR_CUDA;
    const long long tag_ptr = (long long)(size_t)d_dcr;
    if (tag[0] != fft_size || tag[1] != (int)(tag_ptr & 0x7fffffff) || tag[2] != (int)(tag_ptr >> 31)) {
      double *h = (double *)ws->get_pinned("syn_dcr_h", sizeof(double) * fft_size);
      if (!h) return WB_ERR_CUDA;
      std::vector<double> r;
      make_dc_remover(fft_size, r);
      for (int i = 0; i < fft_size; ++i) h[i] = r[i];
      WB_CUDA_CHECK(cudaMemcpyAsync(d_dcr, h, sizeof(double) * fft_size, cudaMemcpyHostToDevice, stream));
      tag[0] = fft_size; tag[1] = (int)(tag_ptr & 0x7fffffff); tag[2] = (int)(tag_ptr >> 31);
    }
  }
  int rc;
  if (!noise_ready) {
    if (rng.wait_skip_in) WB_CUDA_CHECK(cudaStreamWaitEvent(stream, rng.wait_skip_in, 0));
    if ((rc = wb_rng_fill(rng.state, rng.skip_in, d_ncount, (unsigned long long)out_length, d_noise, stream))) return rc;
  }

  int resp_pulses;
  if (f0_upper_bound > 0.0) {
    const double fmax = f0_upper_bound > WB_DEFAULT_F0 ? f0_upper_bound : WB_DEFAULT_F0;
    const double bound = (double)out_length * fmax / fs + 2.0;
    resp_pulses = bound < (double)max_pulses ? (int)bound : max_pulses;
  } else {
    int *h_np = (int *)ws->get_pinned("syn_np_h", sizeof(int));
    if (!h_np) return WB_ERR_CUDA;
    WB_CUDA_CHECK(cudaMemcpyAsync(h_np, d_np, sizeof(int), cudaMemcpyDeviceToHost, stream));
    WB_CUDA_CHECK(cudaStreamSynchronize(stream));
    resp_pulses = *h_np > 0 ? *h_np : 1;
  }
  double *d_resp = (double *)ws->get("syn_resp", sizeof(double) * (size_t)resp_pulses * fft_size);
  if (!d_resp) return WB_ERR_CUDA;

  RespParams p;
  p.sp = d_sp; p.ap = d_ap; p.f0_length = f0_length; p.fs = fs; p.fft_size = fft_size; p.log2n = log2n;
  p.frame_period = frame_period; p.pulse_index = d_pidx; p.pulse_shift = d_pshift; p.vuv = d_vuv;
  p.n_pulses = d_np; p.noise = d_noise; p.dc_remover = d_dcr; p.tw_n = tw_n; p.tw_2n = tw_2n; p.response = d_resp;
  const int binsp = ((fft_size / 2 + 1) + 1) & ~1;
  const size_t smem = sizeof(cplx) * (wb_fft_slots(fft_size / 2) + binsp) + sizeof(double) * (2 * binsp + 128);
  p.max_resp_pulses = resp_pulses;
  p.error_flag = ws->error_flag();
  p.range = nullptr; p.row_begin = 0; p.pulse_vuv = nullptr;
  const int grid = wb_min_i(resp_pulses, wb_sm_count() * 9);
  rc = WB_DISPATCH_LOG2(log2n, 8, 13, {
    if (cudaFuncSetAttribute(response_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WB_ERR_CUDA;
    WB_LAUNCH("response_kernel", response_kernel<L2><<<grid, 256, smem, stream>>>(p));
  });
  if (rc) return rc;
  WB_CUDA_CHECK(cudaGetLastError());
  WB_LAUNCH("ola_kernel", ola_kernel<<<(out_length + 255) / 256, 256, 0, stream>>>(d_resp, d_pidx, d_np, resp_pulses, fft_size, out_length, d_out,
                                                                                  0, out_length, nullptr));
  WB_CUDA_CHECK(cudaGetLastError());
  // (skip_out is not published: Synthesis is the last consumer of the stream in the chain)
  return rng.advance ? wb_rng_advance(rng.state, d_ncount, rng.skip_in, stream) : WB_OK;
}

static int upload_dc_remover(WbWorkspace *ws, int fft_size, double **out, cudaStream_t stream) {
  double *d_dcr = (double *)ws->get("syn_dcr", sizeof(double) * fft_size);
  int *tag = (int *)ws->get_pinned("syn_dcr_tag", sizeof(int) * 4);
  if (!d_dcr || !tag) return WB_ERR_CUDA;
  const long long tag_ptr = (long long)(size_t)d_dcr;
  if (tag[0] != fft_size || tag[1] != (int)(tag_ptr & 0x7fffffff) || tag[2] != (int)(tag_ptr >> 31)) {
    double *h = (double *)ws->get_pinned("syn_dcr_h", sizeof(double) * fft_size);
    if (!h) return WB_ERR_CUDA;
    std::vector<double> r;
    make_dc_remover(fft_size, r);
    for (int i = 0; i < fft_size; ++i) h[i] = r[i];
    WB_CUDA_CHECK(cudaMemcpyAsync(d_dcr, h, sizeof(double) * fft_size, cudaMemcpyHostToDevice, stream));
    tag[0] = fft_size; tag[1] = (int)(tag_ptr & 0x7fffffff); tag[2] = (int)(tag_ptr >> 31);
  }
  *out = d_dcr;
  return WB_OK;
}

// Plan-time tables of the render calls, uploaded on `stream` (callers that render ranges on several streams order
// those streams after this call).
int wb_synthesis_prepare(WbWorkspace *ws, int fft_size, cudaStream_t stream) {
  double *d_dcr = nullptr;
  return upload_dc_remover(ws, fft_size, &d_dcr, stream);
}

// One rank's share of a long stream: the time base and the pulse list of the WHOLE
// stream are on `ws` (every rank computes them from the gathered f0: they are cheap and sequential); this
// renders the pulses that reach into [sample_begin, sample_end) with their whole-stream noise positions and
// overlap-adds them in the reference's order, so the samples equal those of an unsharded run bit for bit.
struct PulseList {   // where the pulse list lives (workspace of a whole-stream time base, or a streaming synthesis)
  const unsigned char *vuv; const unsigned char *pulse_vuv;
  const int *pidx; const double *pshift; const int *np; const unsigned long long *ncount;
  const int *first;   // sample index of the first pulse of the WHOLE stream (the list may hold a window of it)
};
static int render_range_core(WbWorkspace *ws, int fs, int fft_size, double frame_period_ms, int f0_length,
                             const double *d_sp, const double *d_ap, int row_begin, int n_rows, int out_length,
                             int sample_begin, int sample_end, double *d_out, double f0_upper_bound,
                             const WbRngCursor &rng, cudaStream_t stream, const PulseList &pl, int slot = 0,
                             cudaEvent_t rows_ready = nullptr);

int wb_synthesis_render_range(WbWorkspace *ws, int fs, int fft_size, double frame_period_ms, int f0_length,
                              const double *d_sp, const double *d_ap, int row_begin, int n_rows, int out_length,
                              int sample_begin, int sample_end, double *d_out, double f0_upper_bound,
                              const WbRngCursor &rng, cudaStream_t stream, int slot, cudaEvent_t rows_ready) {
  PulseList pl;
  pl.vuv = (unsigned char *)ws->find("syn_vuv"); pl.pulse_vuv = nullptr;
  pl.pidx = (int *)ws->find("syn_pidx"); pl.pshift = (double *)ws->find("syn_pshift");
  pl.np = (int *)ws->find("syn_np"); pl.ncount = (unsigned long long *)ws->find("syn_ncount");
  if (!pl.vuv || !pl.pidx || !pl.pshift || !pl.np || !pl.ncount) return WB_ERR_CUDA;
  pl.first = pl.np + 1;
  return render_range_core(ws, fs, fft_size, frame_period_ms, f0_length, d_sp, d_ap, row_begin, n_rows, out_length,
                           sample_begin, sample_end, d_out, f0_upper_bound, rng, stream, pl, slot, rows_ready);
}

static int render_range_core(WbWorkspace *ws, int fs, int fft_size, double frame_period_ms, int f0_length,
                             const double *d_sp, const double *d_ap, int row_begin, int n_rows, int out_length,
                             int sample_begin, int sample_end, double *d_out, double f0_upper_bound,
                             const WbRngCursor &rng, cudaStream_t stream, const PulseList &pl, int slot,
                             cudaEvent_t rows_ready) {
  if (out_length <= 0) return WB_OK;
  if (sample_begin < 0 || sample_end > out_length || sample_begin > sample_end || row_begin < 0 || n_rows < 0 ||
      row_begin + n_rows > f0_length || !(f0_upper_bound > 0.0))
    return WB_ERR_ARG;
  int log2n = 0;
  while ((1 << log2n) < fft_size) ++log2n;
  if ((1 << log2n) != fft_size || fft_size < 128 || fft_size > 8192) return WB_ERR_UNSUPPORTED;
  const double frame_period = frame_period_ms / 1000.;
  const int n_samples = sample_end - sample_begin;
  const unsigned char *d_vuv = pl.vuv;
  const int *d_pidx = pl.pidx;
  const double *d_pshift = pl.pshift;
  const int *d_np = pl.np;
  const unsigned long long *d_ncount = pl.ncount;
  // The pulses of the range sit in a window of n_samples + fft_size samples, and the noise of the last one
  // runs up to the next pulse: a voiced sample has f0 > lowest_f0 / 2 (interpolation towards an unvoiced
  // frame, synthesis.cpp:225-243), so pulses are less than 2 fs / lowest_f0 < 2 fft_size samples apart.
  const int span = n_samples + 3 * fft_size + 8;
  // (slot: scratch set of the call -- ranges rendered concurrently on different streams use different slots)
  const std::string sfx = slot ? "#" + std::to_string(slot) : std::string();
  double *d_noise = (double *)ws->get("noise_syn_range" + sfx, sizeof(double) * (size_t)span);
  int *d_range = (int *)ws->get("syn_range" + sfx, sizeof(int) * 4);
  unsigned long long *d_npos = (unsigned long long *)ws->get("syn_range_pos" + sfx, sizeof(unsigned long long) * 2);
  double *d_dcr = nullptr;
  int rc = upload_dc_remover(ws, fft_size, &d_dcr, stream);
  if (rc) return rc;
  if (!d_pidx || !d_pshift || !d_np || !d_noise || !d_range || !d_npos) return WB_ERR_CUDA;
  const cplx *tw_n = wb_twiddle_table(fft_size);
  const cplx *tw_2n = wb_twiddle_table(2 * fft_size);
  if (!tw_n || !tw_2n) return WB_ERR_CUDA;
  if (rng.wait_skip_in) WB_CUDA_CHECK(cudaStreamWaitEvent(stream, rng.wait_skip_in, 0));
  WB_LAUNCH("pulse_range_kernel", pulse_range_kernel<<<1, 1, 0, stream>>>(d_pidx, d_np, pl.first, fft_size, sample_begin, sample_end, rng.skip_in,
                                                                     d_range, d_npos, d_npos + 1));
  WB_CUDA_CHECK(cudaGetLastError());
  if (n_samples > 0) {
    if ((rc = wb_rng_fill(rng.state, d_npos, d_npos + 1, (unsigned long long)span, d_noise, stream))) return rc;
    const double fmax = f0_upper_bound > WB_DEFAULT_F0 ? f0_upper_bound : WB_DEFAULT_F0;
    const int resp_pulses = (int)((double)span * fmax / fs + 4.0);
    double *d_resp = (double *)ws->get("syn_resp" + sfx, sizeof(double) * (size_t)resp_pulses * fft_size);
    if (!d_resp) return WB_ERR_CUDA;
    RespParams p;
    p.sp = d_sp; p.ap = d_ap; p.f0_length = f0_length; p.fs = fs; p.fft_size = fft_size; p.log2n = log2n;
    p.frame_period = frame_period; p.pulse_index = d_pidx; p.pulse_shift = d_pshift; p.vuv = d_vuv;
    p.n_pulses = d_np; p.noise = d_noise; p.dc_remover = d_dcr; p.tw_n = tw_n; p.tw_2n = tw_2n; p.response = d_resp;
    const int binsp = ((fft_size / 2 + 1) + 1) & ~1;
    const size_t smem = sizeof(cplx) * (wb_fft_slots(fft_size / 2) + binsp) + sizeof(double) * (2 * binsp + 128);
    p.max_resp_pulses = resp_pulses;
    p.error_flag = ws->error_flag();
    p.range = d_range; p.row_begin = row_begin; p.pulse_vuv = pl.pulse_vuv;
    const int grid = wb_min_i(resp_pulses, wb_sm_count() * 9);
    // (rows_ready: the pulse range and its noise above need the time base only; sp / ap rows may still be landing)
    if (rows_ready) WB_CUDA_CHECK(cudaStreamWaitEvent(stream, rows_ready, 0));
    rc = WB_DISPATCH_LOG2(log2n, 8, 13, {
      if (cudaFuncSetAttribute(response_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WB_ERR_CUDA;
      WB_LAUNCH("response_kernel", response_kernel<L2><<<grid, 256, smem, stream>>>(p));
    });
    if (rc) return rc;
    WB_CUDA_CHECK(cudaGetLastError());
    WB_LAUNCH("ola_kernel", ola_kernel<<<(n_samples + 255) / 256, 256, 0, stream>>>(d_resp, d_pidx, d_np, resp_pulses, fft_size, out_length, d_out,
                                                                                   sample_begin, sample_end, d_range));
    WB_CUDA_CHECK(cudaGetLastError());
  }
  return rng.advance ? wb_rng_advance(rng.state, d_ncount, rng.skip_in, stream) : WB_OK;
}

int wb_synthesis_run(WbWorkspace *ws, int fs, int fft_size, double frame_period_ms, const double *d_f0,
                     int f0_length, const double *d_sp, const double *d_ap, int out_length, double *d_out,
                     double f0_upper_bound, const WbRngCursor &rng, cudaStream_t stream) {
  int rc = wb_synthesis_timebase(ws, fs, fft_size, frame_period_ms, d_f0, f0_length, out_length, stream, nullptr);
  if (rc) return rc;
  return wb_synthesis_render(ws, fs, fft_size, frame_period_ms, f0_length, d_sp, d_ap, out_length, d_out,
                             f0_upper_bound, rng, stream, false);
}

// ---- streaming synthesis ----------------------------------------------------------
// Frames arrive in pieces.  Everything Synthesis::compute derives sequentially is carried between pieces --
// the last phase sum and voicing flag of the time base (synthesis.cpp:257-264), the pulse list, the position
// of the next pulse's noise in the randn() stream (synthesis.cpp:520) -- so the concatenated output equals one
// compute() call on all frames, bit for bit.  Samples are emitted once no later pulse can reach them: a sample n
// is final when every pulse up to n + fft_size/2 - 1 has been rendered, and a pulse can be rendered once its
// successor is known (its noise segment ends there).  A sample t of the time base depends on frames
// floor(t / frame_period) and the next one only, never on the (unknown) end of the contour, as long as the next
// frame is a real one; the extrapolated last point (synthesis.cpp:193-199) is used at finish() only.
struct WbSynStream {
  int fs, fft_size;
  double frame_period_ms, f0_bound;
  WbWorkspace ws;
  int frames = 0;        // frames received
  int row_base = 0;      // first frame held in the sp / ap window
  int tb_done = 0;       // samples whose phase sum is known
  int out_done = 0;      // samples emitted
  int n_pulses = 0;      // pulses found so far
  int final_length = -1; // set by finish()
  bool advanced = false; // randn() state moved past the stream
  std::vector<int> h_pidx;   // host mirror of the pulse sample indices
};

WbSynStream *wb_synstream_create(int fs, int fft_size, double frame_period_ms, double f0_upper_bound) {
  int log2n = 0;
  while ((1 << log2n) < fft_size) ++log2n;
  if (fs <= 0 || (1 << log2n) != fft_size || fft_size < 128 || fft_size > 8192 || !(frame_period_ms > 0.0)) return nullptr;
  WbSynStream *s = new (std::nothrow) WbSynStream();
  if (!s) return nullptr;
  s->fs = fs; s->fft_size = fft_size; s->frame_period_ms = frame_period_ms;
  s->f0_bound = f0_upper_bound > 0.0 ? f0_upper_bound : 1000.0;
  return s;
}
void wb_synstream_destroy(WbSynStream *s) { delete s; }

namespace {
__global__ void synstream_count_kernel(const unsigned long long *__restrict__ block_offsets, int n_blocks, int *__restrict__ np) {
  *np += (int)block_offsets[n_blocks];
}

// time base + exact phase sum + pulses of samples [s->tb_done, n_tb); f0_length = frames seen so far
int synstream_timebase(WbSynStream *s, int n_tb, cudaStream_t stream) {
  const int n_new = n_tb - s->tb_done;
  if (n_new <= 0) return WB_OK;
  WbWorkspace *ws = &s->ws;
  const int off = s->tb_done > 0 ? 1 : 0;   // slot 0 of the local arrays carries the previous piece's last sample
  const int m = n_new + off;
  const double frame_period = s->frame_period_ms / 1000.;
  const double lowest_f0 = s->fs / s->fft_size + 1.0;             // synthesis.cpp:97 (integer division)
  double *d_f0 = (double *)ws->find("st_f0");
  double *d_incr = (double *)ws->get("st_incr", sizeof(double) * m);
  double *d_total = (double *)ws->get("st_total", sizeof(double) * m);
  unsigned char *d_vuv = (unsigned char *)ws->get("st_vuv", m);
  double *d_carry = (double *)ws->get("st_carry", 16);             // [0] last phase sum, [1] (first byte) last voicing flag
  const int n_blocks = (m + PD_THREADS - 1) / PD_THREADS;
  unsigned long long *d_bc = (unsigned long long *)ws->get("syn_bcount", sizeof(unsigned long long) * (n_blocks + 1));
  unsigned long long *d_bo = (unsigned long long *)ws->get("syn_boff", sizeof(unsigned long long) * (n_blocks + 1));
  int *d_np = (int *)ws->get("st_np", sizeof(int) * 4);
  // a piece of m samples holds at most m * f0_bound / fs + 2 pulses
  const int cap_new = (int)((double)m * (s->f0_bound > WB_DEFAULT_F0 ? s->f0_bound : WB_DEFAULT_F0) / s->fs) + 4;
  const int cap = s->n_pulses + cap_new;
  int *d_pidx = (int *)ws->get_keep("st_pidx", sizeof(int) * cap, sizeof(int) * s->n_pulses, stream);
  double *d_pshift = (double *)ws->get_keep("st_pshift", sizeof(double) * cap, sizeof(double) * s->n_pulses, stream);
  unsigned char *d_pvuv = (unsigned char *)ws->get_keep("st_pvuv", cap, s->n_pulses, stream);
  if (!d_f0 || !d_incr || !d_total || !d_vuv || !d_carry || !d_bc || !d_bo || !d_np || !d_pidx || !d_pshift || !d_pvuv) return WB_ERR_CUDA;
  if (off) {
    WB_CUDA_CHECK(cudaMemcpyAsync(d_incr, d_carry, sizeof(double), cudaMemcpyDeviceToDevice, stream));
    WB_CUDA_CHECK(cudaMemcpyAsync(d_vuv, d_carry + 1, 1, cudaMemcpyDeviceToDevice, stream));
  }
  WB_LAUNCH("timebase_kernel", timebase_kernel<<<(n_new + 255) / 256, 256, 0, stream>>>(
      d_f0, s->frames, s->fs, frame_period, lowest_f0, n_tb, d_incr, d_vuv, s->tb_done, off));
  int rc = run_phase_scan(ws, d_incr, m, d_total, stream);
  if (rc) return rc;
  WB_CUDA_CHECK(cudaMemcpyAsync(d_carry, d_total + (m - 1), sizeof(double), cudaMemcpyDeviceToDevice, stream));
  WB_CUDA_CHECK(cudaMemcpyAsync(d_carry + 1, d_vuv + (m - 1), 1, cudaMemcpyDeviceToDevice, stream));
  WB_LAUNCH("pulse_count_kernel", pulse_count_kernel<<<n_blocks, PD_THREADS, 0, stream>>>(d_total, m, d_bc));
  if ((rc = wb_exclusive_scan_u64(d_bc, d_bo, n_blocks, stream))) return rc;
  WB_LAUNCH("pulse_write_kernel", pulse_write_kernel<<<n_blocks, PD_THREADS, 0, stream>>>(d_total, m, s->fs, d_bo, d_pidx, d_pshift, cap,
                                                                                      s->tb_done - off, s->n_pulses, d_vuv, d_pvuv));
  WB_LAUNCH("synstream_count_kernel", synstream_count_kernel<<<1, 1, 0, stream>>>(d_bo, n_blocks, d_np));
  WB_CUDA_CHECK(cudaGetLastError());
  int h_np = 0;
  WB_CUDA_CHECK(cudaMemcpyAsync(&h_np, d_np, sizeof(int), cudaMemcpyDeviceToHost, stream));
  WB_CUDA_CHECK(cudaStreamSynchronize(stream));
  if (h_np < s->n_pulses || h_np > cap) return WB_ERR_ARG;   // f0 above the bound given at creation
  if (h_np > s->n_pulses) {
    s->h_pidx.resize(h_np);
    WB_CUDA_CHECK(cudaMemcpyAsync(s->h_pidx.data() + s->n_pulses, d_pidx + s->n_pulses, sizeof(int) * (h_np - s->n_pulses),
                                  cudaMemcpyDeviceToHost, stream));
    WB_CUDA_CHECK(cudaStreamSynchronize(stream));
  }
  s->n_pulses = h_np;
  s->tb_done = n_tb;
  return WB_OK;
}

// renders and emits samples [s->out_done, out_new)
int synstream_emit(WbSynStream *s, int out_new, int out_length_for_ola, double *out, cudaStream_t stream) {
  const int count = out_new - s->out_done;
  if (count <= 0) return WB_OK;
  WbWorkspace *ws = &s->ws;
  const int bins = s->fft_size / 2 + 1;
  double *d_out = (double *)ws->get("st_out", sizeof(double) * count);
  if (!d_out) return WB_ERR_CUDA;
  PulseList pl;
  pl.vuv = nullptr; pl.pulse_vuv = (unsigned char *)ws->find("st_pvuv");
  pl.pidx = (int *)ws->find("st_pidx"); pl.pshift = (double *)ws->find("st_pshift");
  pl.np = (int *)ws->find("st_np"); pl.ncount = nullptr; pl.first = pl.pidx;   // (the list starts at the stream's first pulse)
  WbRngCursor c;
  c.state = wb_rng_global_state(); c.advance = false;   // the state moves once, when the stream is finished
  const double *d_sp = (const double *)ws->find("st_sp"), *d_ap = (const double *)ws->find("st_ap");
  int rc = render_range_core(ws, s->fs, s->fft_size, s->frame_period_ms, s->frames, d_sp, d_ap, s->row_base,
                             s->frames - s->row_base, out_length_for_ola, s->out_done, out_new, d_out, s->f0_bound, c, stream, pl);
  if (rc) return rc;
  WB_CUDA_CHECK(cudaMemcpyAsync(out, d_out, sizeof(double) * count, cudaMemcpyDeviceToHost, stream));
  WB_CUDA_CHECK(cudaStreamSynchronize(stream));
  if ((rc = ws->read_error_flag(stream))) return rc;
  s->out_done = out_new;
  // rows no future pulse interpolates any more (a pulse reaching into sample n sits at n - fft_size/2 or later)
  const int keep_from = wb_max_i(0, (int)floor((double)(s->out_done - s->fft_size) / s->fs / (s->frame_period_ms / 1000.)) - 2);
  if (keep_from > s->row_base) {
    const size_t kept = (size_t)(s->frames - keep_from) * bins;
    if (kept > 0) {
      double *tmp = (double *)ws->get("st_rows_tmp", sizeof(double) * kept);
      double *sp = (double *)ws->find("st_sp"), *ap = (double *)ws->find("st_ap");
      if (!tmp) return WB_ERR_CUDA;
      const size_t shift = (size_t)(keep_from - s->row_base) * bins;
      WB_CUDA_CHECK(cudaMemcpyAsync(tmp, sp + shift, sizeof(double) * kept, cudaMemcpyDeviceToDevice, stream));
      WB_CUDA_CHECK(cudaMemcpyAsync(sp, tmp, sizeof(double) * kept, cudaMemcpyDeviceToDevice, stream));
      WB_CUDA_CHECK(cudaMemcpyAsync(tmp, ap + shift, sizeof(double) * kept, cudaMemcpyDeviceToDevice, stream));
      WB_CUDA_CHECK(cudaMemcpyAsync(ap, tmp, sizeof(double) * kept, cudaMemcpyDeviceToDevice, stream));
    }
    s->row_base = keep_from;
  }
  return WB_OK;
}
}  // namespace

int wb_synstream_push(WbSynStream *s, const double *f0, const double *sp, const double *ap, int n_frames, double *out,
                      int out_capacity, int *n_out, cudaStream_t stream) {
  if (!s || !n_out || n_frames < 0 || out_capacity < 0 || (n_frames > 0 && (!f0 || !sp || !ap)) || (out_capacity > 0 && !out) ||
      s->final_length >= 0)
    return WB_ERR_ARG;
  *n_out = 0;
  WbWorkspace *ws = &s->ws;
  const size_t bins = s->fft_size / 2 + 1;
  if (n_frames > 0) {
    const int total = s->frames + n_frames, held = s->frames - s->row_base;
    double *d_f0 = (double *)ws->get_keep("st_f0", sizeof(double) * total, sizeof(double) * s->frames, stream);
    double *d_sp = (double *)ws->get_keep("st_sp", sizeof(double) * (held + n_frames) * bins, sizeof(double) * held * bins, stream);
    double *d_ap = (double *)ws->get_keep("st_ap", sizeof(double) * (held + n_frames) * bins, sizeof(double) * held * bins, stream);
    int *d_np = (int *)ws->get("st_np", sizeof(int) * 4);
    if (!d_f0 || !d_sp || !d_ap || !d_np) return WB_ERR_CUDA;
    if (s->frames == 0) WB_CUDA_CHECK(cudaMemsetAsync(d_np, 0, sizeof(int) * 4, stream));
    WB_CUDA_CHECK(cudaMemcpyAsync(d_f0 + s->frames, f0, sizeof(double) * n_frames, cudaMemcpyHostToDevice, stream));
    WB_CUDA_CHECK(cudaMemcpyAsync(d_sp + (size_t)held * bins, sp, sizeof(double) * n_frames * bins, cudaMemcpyHostToDevice, stream));
    WB_CUDA_CHECK(cudaMemcpyAsync(d_ap + (size_t)held * bins, ap, sizeof(double) * n_frames * bins, cudaMemcpyHostToDevice, stream));
    WB_CUDA_CHECK(cudaStreamSynchronize(stream));   // the caller's buffers may be reused on return
    s->frames = total;
  }
  if (s->frames < 2) return WB_OK;
  // samples whose two frames are both real ones: t = ii / fs < (frames - 1) * frame_period, with the kernel's own expressions
  const double frame_period = s->frame_period_ms / 1000.;
  const double t_end = (s->frames - 1) * frame_period;
  long long n_tb = (long long)(t_end * s->fs) - 2;
  if (n_tb < 0) n_tb = 0;
  while (n_tb < 2147483000LL && (double)n_tb / (double)s->fs < t_end) ++n_tb;
  int rc = synstream_timebase(s, (int)n_tb, stream);
  if (rc) return rc;
  if (s->n_pulses < 2) return WB_OK;
  // every pulse but the last known one can be rendered; samples up to (last known pulse) - fft_size/2 are final
  int out_new = s->h_pidx[s->n_pulses - 1] - s->fft_size / 2 + 1;
  if (out_new > s->out_done + out_capacity) out_new = s->out_done + out_capacity;
  if (out_new <= s->out_done) return WB_OK;
  const int before = s->out_done;
  if ((rc = synstream_emit(s, out_new, 2147483647, out, stream))) return rc;
  *n_out = s->out_done - before;
  return WB_OK;
}

int wb_synstream_finish(WbSynStream *s, int out_length_total, double *out, int out_capacity, int *n_out, cudaStream_t stream) {
  if (!s || !n_out || out_capacity < 0 || (out_capacity > 0 && !out) || s->frames < 2) return WB_ERR_ARG;
  *n_out = 0;
  if (s->final_length < 0) {
    if (out_length_total < s->tb_done || out_length_total < s->out_done) return WB_ERR_ARG;   // shorter than what already went out
    s->final_length = out_length_total;
    int rc = synstream_timebase(s, out_length_total, stream);   // the rest, with the extrapolated last point
    if (rc) return rc;
  } else if (out_length_total != s->final_length) {
    return WB_ERR_ARG;
  }
  int out_new = s->final_length;
  if (out_new > s->out_done + out_capacity) out_new = s->out_done + out_capacity;
  const int before = s->out_done;
  int rc = synstream_emit(s, out_new, s->final_length, out, stream);
  if (rc) return rc;
  *n_out = s->out_done - before;
  if (s->out_done == s->final_length && !s->advanced) {
    // leave the randn() state where one compute() call would have left it: noise_size summed over the pulses
    unsigned long long *h = (unsigned long long *)s->ws.get_pinned("st_ncount_h", sizeof(unsigned long long));
    unsigned long long *d = (unsigned long long *)s->ws.get("st_ncount", sizeof(unsigned long long));
    if (!h || !d) return WB_ERR_CUDA;
    *h = s->n_pulses >= 2 ? (unsigned long long)(s->h_pidx[s->n_pulses - 1] - s->h_pidx[0]) : 0ull;
    WB_CUDA_CHECK(cudaMemcpyAsync(d, h, sizeof(*h), cudaMemcpyHostToDevice, stream));
    if ((rc = wb_rng_advance(wb_rng_global_state(), d, nullptr, stream))) return rc;
    WB_CUDA_CHECK(cudaStreamSynchronize(stream));
    s->advanced = true;
  }
  return WB_OK;
}
