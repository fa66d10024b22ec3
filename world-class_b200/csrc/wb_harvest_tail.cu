// Harvest, part 2: from the pruned candidate table to the smoothed F0 contour.
//
// Reference: /root/reference/src/harvest.cpp
//   searchF0Base :254-272, fixStep1 :277-291, getBoundaryList :296-314, fixStep2 :319-334,
//   selectBestF0 :347-365, extendF0 :371-404, swapArray :410-424, extend :429-458,
//   searchScore :463-470, mergeF0Sub :475-497, mergeF0 :502-536, getMultiChannelF0 :542-555,
//   fixStep3 :560-585, fixStep4 :590-614, fixF0Contour :619-634, filteringF0 :639-665,
//   smoothF0Contour :670-703.
//
// This is branchy, mostly serial bookkeeping over a few dozen voiced sections.  It runs as
// ONE thread block so that the whole Harvest stage stays on the device with no host round
// trip: per-frame work is spread over the 1024 threads, per-section work over warps, and the
// two zero-phase Butterworth passes are evaluated in parallel chunks with a 300-sample
// warm-up (pole radius 0.875 -> 4e-18 residual), which is also the padding the reference uses.
//
// Voiced sections are stored sparsely (section range +-104 frames) instead of the reference's
// f0_length-sized row per section; reads outside the stored range return the 0.0 the
// reference's rows hold there.
#include "wb_harvest.h"

#include <algorithm>

namespace {

#define TL_THREADS 512
#define TL_MARGIN 104
#define TL_LAG 300
#define TL_CHUNK 31   // odd: the per-thread stride stays bank-conflict free in shared memory

struct TailParams {
  const double *cand; const double *score; const int *nc; int L; int MC;
  double *base, *s1, *s2, *s3, *s4, *out;
  int *blist;          // L + 2*TL_LAG + 2
  int maxsec;
  int *sec_st, *sec_ed, *sec_lo, *sec_len, *sec_off;  // maxsec each (sec_off: maxsec + 1)
  double *sec_sum;     // maxsec
  double *secbuf; long long secbuf_cap;
  int *kb;             // 2 * maxsec (+2)
  int *kslot;          // maxsec (+1)
  int *order;          // maxsec
  double *pad;         // L + 2*TL_LAG
  double *fw; long long fw_cap;
  double *gA, *gB;     // global fallbacks of the two contour buffers (L + 2*TL_LAG each)
  int smem_doubles;    // dynamic shared memory available to the kernel
  long long *clocks;   // [16] phase time stamps (clock64 of thread 0), for profiling the serial tail
  int *error_flag;
  int *smooth_hdr;     // [4] section count / forward items / backward items of the smoothing passes
};

__device__ __forceinline__ int tl_vuv(const double *f0, int n, int i) {
  return (i <= 0 || i >= n - 1) ? 0 : (f0[i] > 0 ? 1 : 0);
}

// getBoundaryList (harvest.cpp:296-314).  All threads; returns the number of boundaries.  The list goes to
// the shared-memory buffer when it fits (the usual case: a few dozen voiced sections), else to global
// scratch; *blist_out tells which.
__device__ int tl_boundaries(const double *f0, int n, int *blist_shared, int shared_cap, int *blist_global,
                             int **blist_out, int *s_scan) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int chunk = (n + nt - 1) / nt;
  const int b = max(1, tid * chunk), e = min(n, (tid + 1) * chunk);
  int cnt = 0;
  for (int i = b; i < e; ++i) cnt += (tl_vuv(f0, n, i) - tl_vuv(f0, n, i - 1) != 0) ? 1 : 0;
  __syncthreads();
  s_scan[tid] = cnt;
  __syncthreads();
  for (int o = 1; o < nt; o <<= 1) {
    const int t = (tid >= o) ? s_scan[tid - o] : 0;
    __syncthreads();
    s_scan[tid] += t;
    __syncthreads();
  }
  int k = s_scan[tid] - cnt;
  const int total = s_scan[nt - 1];
  int *blist = (total <= shared_cap) ? blist_shared : blist_global;
  *blist_out = blist;
  for (int i = b; i < e; ++i) {
    if (tl_vuv(f0, n, i) - tl_vuv(f0, n, i - 1) != 0) { blist[k] = i - k % 2; ++k; }
  }
  __syncthreads();
  return total;
}

__device__ __forceinline__ double tl_getsec(const TailParams &p, int s, int i) {
  const int lo = p.sec_lo[s];
  const int j = i - lo;
  return (j >= 0 && j < p.sec_len[s]) ? p.secbuf[p.sec_off[s] + j] : 0.0;
}

// selectBestF0 (harvest.cpp:347-365) by one warp: minimum error, ties -> the LAST candidate
__device__ __forceinline__ double tl_select_best(double reference_f0, const double *cands, int n, double allowed) {
  const int lane = threadIdx.x & 31;
  double best_err = allowed;
  int best_idx = -1;
  for (int i = lane; i < n; i += 32) {
    const double tmp = fabs(reference_f0 - cands[i]) / reference_f0;
    if (tmp > best_err) continue;
    best_err = tmp;
    best_idx = i;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double e2 = __shfl_xor_sync(0xffffffffu, best_err, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, best_idx, o);
    // valid entries only (idx >= 0); smaller error wins, equal error -> larger index
    if (i2 >= 0 && (best_idx < 0 || e2 < best_err || (e2 == best_err && i2 > best_idx))) { best_err = e2; best_idx = i2; }
  }
  return best_idx >= 0 ? cands[best_idx] : 0.0;
}

// selectBestF0 on a candidate row held in registers (lane l: entries l and l + 32), resolved with
// warp reductions instead of a shuffle tree: fl(d / ref) is monotone in d, so the candidate with the
// smallest |ref - c| (the LAST one among equal distances) is the reference's choice, and one
// division at the end reproduces its `tmp > best_error` test against allowed_range.
__device__ __forceinline__ double tl_select_best_fast(double reference_f0, double c0, double c1, int n, double allowed) {
  const int lane = threadIdx.x & 31;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  const double d0 = (lane < n) ? fabs(reference_f0 - c0) : inf;
  const double d1 = (lane + 32 < n) ? fabs(reference_f0 - c1) : inf;
  double dl = d0, vl = c0;
  int il = lane;
  if (lane + 32 < n && d1 <= d0) { dl = d1; vl = c1; il = lane + 32; }
  const unsigned long long key = (unsigned long long)__double_as_longlong(dl);  // dl >= 0 or +inf (NaN sorts last)
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
  const int mi = __reduce_max_sync(0xffffffffu, (hi == mh && lo == ml) ? il : -1);
  if (mi < 0) return 0.0;
  const double v = __shfl_sync(0xffffffffu, vl, mi & 31);
  const double dmin = __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
  // `fl(dmin / ref) > allowed`: the quotient (the long pole of this serial chain) is only formed when the
  // decision is within rounding distance of the threshold
  const double bound = allowed * reference_f0;
  if (dmin < bound * (1.0 - 1e-12)) return v;
  if (dmin > bound * (1.0 + 1e-12)) return 0.0;
  const double err = dmin / reference_f0;
  return (err > allowed) ? 0.0 : v;
}

// (shuffle-tree version, kept for reference / wide tables)
__device__ __forceinline__ double tl_select_best_reg(double reference_f0, double c0, double c1, int n, double allowed) {
  const int lane = threadIdx.x & 31;
  double best_err = allowed, best_val = 0.0;
  int best_idx = -1;
  if (lane < n) {
    const double tmp = fabs(reference_f0 - c0) / reference_f0;
    if (!(tmp > best_err)) { best_err = tmp; best_idx = lane; best_val = c0; }
  }
  if (lane + 32 < n) {
    const double tmp = fabs(reference_f0 - c1) / reference_f0;
    if (!(tmp > best_err)) { best_err = tmp; best_idx = lane + 32; best_val = c1; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double e2 = __shfl_xor_sync(0xffffffffu, best_err, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, best_idx, o);
    const double v2 = __shfl_xor_sync(0xffffffffu, best_val, o);
    if (i2 >= 0 && (best_idx < 0 || e2 < best_err || (e2 == best_err && i2 > best_idx))) { best_err = e2; best_idx = i2; best_val = v2; }
  }
  return best_idx >= 0 ? best_val : 0.0;
}

// extendF0 (harvest.cpp:371-404) by one warp on sparse section s.  The candidate rows of the next
// TL_PF frames are fetched together (they do not depend on the running F0), so the serial chain
// of selectBestF0 calls runs from registers instead of paying one L2 round trip per frame.
#define TL_PF 8
__device__ int tl_extend_f0(const TailParams &p, int s, int origin, int last_point, int shift, int nc7) {
  const int lane = threadIdx.x & 31;
  double *f = p.secbuf + p.sec_off[s] - p.sec_lo[s];  // f[i] valid for lo <= i <= hi
  const int threshold = 4;
  double tmp_f0 = f[origin];
  int shifted_origin = origin;
  const int distance = abs(last_point - origin);
  int count = 0;
  bool stop = false;
  const int n = nc7;
  if (n > 64) {
    // wide candidate tables (rare): plain version, one row per step from global memory
    for (int i = 0; i <= distance; ++i) {
      const int pos = origin + shift * i + shift;
      const double v = tl_select_best(tmp_f0, p.cand + (size_t)pos * p.MC, nc7, 0.18);
      if (lane == 0) f[pos] = v;
      if (v == 0.0) {
        count++;
      } else {
        tmp_f0 = v;
        count = 0;
        shifted_origin = pos;
      }
      if (count == threshold) break;
    }
    __syncwarp();
    return shifted_origin;
  }
  // software pipeline: the rows of batch k+1 are in flight while batch k is resolved
  double n0[TL_PF], n1[TL_PF];
  auto load_batch = [&](int i0, double (&c0)[TL_PF], double (&c1)[TL_PF]) {
#pragma unroll
    for (int k = 0; k < TL_PF; ++k) {
      const int i = i0 + k;
      c0[k] = 0.0; c1[k] = 0.0;
      if (i <= distance) {
        const double *row = p.cand + (size_t)(origin + shift * i + shift) * p.MC;
        if (lane < n) c0[k] = row[lane];
        if (lane + 32 < n) c1[k] = row[lane + 32];
      }
    }
  };
  load_batch(0, n0, n1);
  for (int i0 = 0; i0 <= distance && !stop; i0 += TL_PF) {
    double c0[TL_PF], c1[TL_PF];
#pragma unroll
    for (int k = 0; k < TL_PF; ++k) { c0[k] = n0[k]; c1[k] = n1[k]; }
    if (i0 + TL_PF <= distance) load_batch(i0 + TL_PF, n0, n1);
#pragma unroll
    for (int k = 0; k < TL_PF; ++k) {
      const int i = i0 + k;
      if (i <= distance && !stop) {
        const int pos = origin + shift * i + shift;
        const double v = tl_select_best_fast(tmp_f0, c0[k], c1[k], n, 0.18);
        if (lane == 0) f[pos] = v;
        if (v == 0.0) {
          count++;
        } else {
          tmp_f0 = v;
          count = 0;
          shifted_origin = pos;
        }
        if (count == threshold) stop = true;
      }
    }
  }
  __syncwarp();
  return shifted_origin;
}

// searchScore (harvest.cpp:463-470)
__device__ __forceinline__ double tl_search_score(double f0, const double *cands, const double *scores, int n) {
  double score = 0.0;
  for (int i = 0; i < n; ++i)
    if (f0 == cands[i] && score < scores[i]) score = scores[i];
  return score;
}

// searchF0Base (harvest.cpp:254-272): one warp per frame; the first candidate with the highest
// positive score wins (the reference's scan uses a strict >).
__global__ void search_base_kernel(const double *__restrict__ cand, const double *__restrict__ score,
                                   const int *__restrict__ nc, int L, int MC, double *__restrict__ base) {
  const int frame = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (frame >= L) return;
  const int nc7 = *nc * 7;
  const double *c = cand + (size_t)frame * MC, *sc = score + (size_t)frame * MC;
  double best_score = 0.0;
  int best_idx = -1;
  for (int j = lane; j < nc7; j += 32) {
    const double v = sc[j];
    if (v > best_score) { best_score = v; best_idx = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double s2 = __shfl_xor_sync(0xffffffffu, best_score, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, best_idx, o);
    if (i2 >= 0 && (best_idx < 0 || s2 > best_score || (s2 == best_score && i2 < best_idx))) { best_score = s2; best_idx = i2; }
  }
  if (lane == 0) base[frame] = best_idx >= 0 ? c[best_idx] : 0.0;
}

// exclusive prefix sum of v[0..n) into out[0..n] (out[n] = total); all threads; n may exceed blockDim
__device__ void tl_exclusive_scan(const int *v, int n, int *out, int *s_scan) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int chunk = (n + nt - 1) / nt;
  const int b = min(n, tid * chunk), e = min(n, (tid + 1) * chunk);
  int cnt = 0;
  for (int i = b; i < e; ++i) cnt += v[i];
  __syncthreads();
  s_scan[tid] = cnt;
  __syncthreads();
  for (int o = 1; o < nt; o <<= 1) {
    const int t = (tid >= o) ? s_scan[tid - o] : 0;
    __syncthreads();
    s_scan[tid] += t;
    __syncthreads();
  }
  int run = s_scan[tid] - cnt;
  for (int i = b; i < e; ++i) { const int x = v[i]; out[i] = run; run += x; }
  if (tid == nt - 1) out[n] = s_scan[nt - 1];
  __syncthreads();
}

#define TL_SMAX 192   // sections whose bookkeeping fits in shared memory
__global__ void __launch_bounds__(TL_THREADS) harvest_tail_kernel(TailParams p_in) {
  extern __shared__ double tl_smem[];   // p.smem_doubles doubles: contour buffers A | B (later: forward-filter scratch)
  __shared__ int s_scan[TL_THREADS];
  // Section bookkeeping (boundaries, ranges, offsets, ...) is read in chains of dependent loads by a few
  // threads: it lives in shared memory whenever the number of sections allows (else in global scratch).
  __shared__ int s_blist[2 * TL_SMAX + 2];
  __shared__ int s_secint[9 * (TL_SMAX + 2)];
  __shared__ double s_secsum[TL_SMAX + 2];
  TailParams p = p_in;
  auto use_shared_sections = [&](bool yes) {
    if (yes) {
      const int m = TL_SMAX + 2;
      p.sec_st = s_secint; p.sec_ed = s_secint + m; p.sec_lo = s_secint + 2 * m; p.sec_len = s_secint + 3 * m;
      p.sec_off = s_secint + 4 * m; p.kb = s_secint + 5 * m; p.kslot = s_secint + 7 * m; p.order = s_secint + 8 * m;
      p.sec_sum = s_secsum;
    } else {
      p.sec_st = p_in.sec_st; p.sec_ed = p_in.sec_ed; p.sec_lo = p_in.sec_lo; p.sec_len = p_in.sec_len;
      p.sec_off = p_in.sec_off; p.kb = p_in.kb; p.kslot = p_in.kslot; p.order = p_in.order; p.sec_sum = p_in.sec_sum;
    }
  };
  __shared__ double s_red[64];
  __shared__ int s_i[8];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const int L = p.L, MC = p.MC;
  const int nc7 = *p.nc * 7;
  const int Lp_cap = L + 2 * TL_LAG;
  // Two contour buffers A and B (padded length each) live in shared memory when they fit (10 s of
  // audio at 1 ms: 2 x 85 KB), otherwise in global scratch.  During smoothing only A (the padded
  // contour) is live and everything above it is forward-filter scratch.
  const bool two_ok = p.smem_doubles >= 2 * Lp_cap;
  double *cA = two_ok ? tl_smem : p.gA;
  double *cB = two_ok ? tl_smem + Lp_cap : p.gB;
  // (the smoothing kernels that follow read this header: no work unless the set-up at the end fills it in)
  if (tid == 0) { p_in.smooth_hdr[0] = 0; p_in.smooth_hdr[1] = 0; p_in.smooth_hdr[2] = 0; }
  int clk_i = 0;
#define TL_STAMP() do { if (tid == 0) p.clocks[clk_i] = clock64(); ++clk_i; } while (0)
  TL_STAMP();  // 0: start

  // ---- fixStep1 (searchF0Base ran in search_base_kernel)
  for (int i = tid; i < L; i += nt) cA[i] = p.base[i];
  __syncthreads();
  for (int i = tid; i < L; i += nt) {
    double v = 0.0;  // entries the reference leaves unwritten read as 0 (zero-filled heap)
    const double b0v = cA[i];
    if (i >= 2 && b0v != 0.0) {
      const double b1v = cA[i - 1], b2v = cA[i - 2];
      const double reference_f0 = b1v * 2 - b2v;
      v = (fabs((b0v - reference_f0) / reference_f0) > 0.008 && fabs((b0v - b1v)) / b1v > 0.008) ? 0.0 : b0v;
    }
    cB[i] = v;
    p.s1[i] = v;
  }
  __syncthreads();

  TL_STAMP();  // 1: after fixStep1
  // ---- fixStep2: drop voiced sections shorter than 6 (in place: the boundary list is extracted first)
  {
    const int nb = tl_boundaries(cB, L, s_blist, 2 * TL_SMAX, p_in.blist, &p.blist, s_scan);
    for (int k = tid; k < nb / 2; k += nt) {
      const int st = p.blist[2 * k], ed = p.blist[2 * k + 1];
      if (ed - st >= 6) continue;
      for (int j = st; j <= ed; ++j) cB[j] = 0.0;
    }
    __syncthreads();
    for (int i = tid; i < L; i += nt) p.s2[i] = cB[i];
  }

  TL_STAMP();  // 2: after fixStep2
  // ---- fixStep3
  int nsec;
  {
    const int nb = tl_boundaries(cB, L, s_blist, 2 * TL_SMAX, p_in.blist, &p.blist, s_scan);
    nsec = nb / 2;
    use_shared_sections(nsec <= TL_SMAX);
    if (nsec > p.maxsec) {
      if (tid == 0) atomicExch(p.error_flag, WB_ERR_UNSUPPORTED);
      nsec = 0;
    }
  }
  if (nsec == 0) {
    // no voiced section: the reference reads multi_channel_f0[0] of an empty array (undefined);
    // we define the result as the all-unvoiced contour.
    for (int i = tid; i < L; i += nt) p.out[i] = 0.0;
    return;
  }
  for (int s = tid; s < nsec; s += nt) {
    const int st = p.blist[2 * s], ed = p.blist[2 * s + 1];
    const int lo = max(0, st - TL_MARGIN), hi = min(L - 1, ed + TL_MARGIN);
    p.sec_st[s] = st; p.sec_ed[s] = ed; p.sec_lo[s] = lo; p.sec_len[s] = hi - lo + 1;
  }
  __syncthreads();
  tl_exclusive_scan(p.sec_len, nsec, p.sec_off, s_scan);
  if (tid == 0) s_i[0] = ((long long)p.sec_off[nsec] > p.secbuf_cap) ? 1 : 0;
  __syncthreads();
  if (s_i[0]) {
    if (tid == 0) atomicExch(p.error_flag, WB_ERR_UNSUPPORTED);
    for (int i = tid; i < L; i += nt) p.out[i] = 0.0;
    return;
  }
  // getMultiChannelF0
  for (int s = 0; s < nsec; ++s) {
    const int st = p.sec_st[s], ed = p.sec_ed[s], lo = p.sec_lo[s], len = p.sec_len[s], off = p.sec_off[s];
    for (int j = tid; j < len; j += nt) {
      const int i = lo + j;
      p.secbuf[off + j] = (i >= st && i <= ed) ? cB[i] : 0.0;
    }
  }
  __syncthreads();
  TL_STAMP();  // 3: sections materialised
  // extend: the forward and the backward extension of a section are independent serial chains (they write
  // disjoint frames of the section and start from its own two ends), so they run on different warps;
  // then the per-section sums for extendSub
  for (int item = warp; item < 2 * nsec; item += nwarps) {
    const int s = item >> 1;
    if ((item & 1) == 0) {
      const int ed0 = p.sec_ed[s];
      const int ed1 = tl_extend_f0(p, s, ed0, min(L - 2, ed0 + 100), 1, nc7);
      if (lane == 0) p.kb[2 * s + 1] = ed1;
    } else {
      const int st0 = p.sec_st[s];
      const int st1 = tl_extend_f0(p, s, st0, max(1, st0 - 100), -1, nc7);
      if (lane == 0) p.kb[2 * s] = st1;
    }
  }
  __syncthreads();
  for (int s = warp; s < nsec; s += nwarps) {
    const int st1 = p.kb[2 * s], ed1 = p.kb[2 * s + 1];
    const double *f = p.secbuf + p.sec_off[s] - p.sec_lo[s];
    double acc = 0.0;
    for (int j = st1 + lane; j < ed1; j += 32) acc += f[j];
    acc = wb_warp_sum(acc);
    if (lane == 0) { p.sec_st[s] = st1; p.sec_ed[s] = ed1; p.sec_sum[s] = acc; }
  }
  __syncthreads();
  TL_STAMP();  // 4: after extend
  // extendSub (harvest.cpp:443-455); mean_f0 is carried across sections
  if (tid == 0) {
    const double threshold2 = 2200.0;
    int count = 0;
    double mean_f0 = 0.0;
    for (int s = 0; s < nsec; ++s) {
      const int st = p.sec_st[s], ed = p.sec_ed[s];
      mean_f0 += p.sec_sum[s];
      mean_f0 /= ed - st;
      if (threshold2 / mean_f0 < ed - st) {
        p.kslot[count] = s; p.kb[2 * count] = st; p.kb[2 * count + 1] = ed;
        count++;
      }
    }
    if (count == 0) { p.kslot[0] = 0; p.kb[0] = p.sec_st[0]; p.kb[1] = p.sec_ed[0]; }
    s_i[1] = count;
    // argsort by start (harvest.cpp:510-514); stable insertion sort
    for (int i = 0; i < count; ++i) p.order[i] = i;
    for (int i = 1; i < count; ++i) {
      const int v = p.order[i];
      int j = i - 1;
      while (j >= 0 && p.kb[2 * p.order[j]] > p.kb[2 * v]) { p.order[j + 1] = p.order[j]; --j; }
      p.order[j + 1] = v;
    }
  }
  __syncthreads();
  const int nkeep = s_i[1];
  TL_STAMP();  // 5: after extendSub + sort
  // mergeF0 (harvest.cpp:502-536)
  {
    const int s0 = p.kslot[0];
    for (int i = tid; i < L; i += nt) cA[i] = tl_getsec(p, s0, i);
    __syncthreads();
    for (int it = 1; it < nkeep; ++it) {
      const int o = p.order[it];
      const int so = p.kslot[o];
      const int index1 = p.kb[2 * o], index2 = p.kb[2 * o + 1];
      const int st1 = p.kb[0], ed1 = p.kb[1];
      int new_b0 = st1, new_b1 = ed1;
      if (index1 - ed1 > 0) {
        for (int i = index1 + tid; i <= index2; i += nt) cA[i] = tl_getsec(p, so, i);
        new_b0 = index1; new_b1 = index2;
      } else if (!(st1 <= index1 && ed1 >= index2)) {
        // mergeF0Sub
        double sc1 = 0.0, sc2 = 0.0;
        for (int i = index1 + tid; i <= ed1; i += nt) {
          const double *c = p.cand + (size_t)i * MC, *sc = p.score + (size_t)i * MC;
          sc1 += tl_search_score(cA[i], c, sc, nc7);
          sc2 += tl_search_score(tl_getsec(p, so, i), c, sc, nc7);
        }
        wb_block_sum2(sc1, sc2, s_red);
        __syncthreads();
        if (sc1 > sc2) { for (int i = ed1 + tid; i <= index2; i += nt) cA[i] = tl_getsec(p, so, i); }
        else { for (int i = index1 + tid; i <= index2; i += nt) cA[i] = tl_getsec(p, so, i); }
        new_b1 = index2;
      }
      __syncthreads();
      if (tid == 0) { p.kb[0] = new_b0; p.kb[1] = new_b1; }
      __syncthreads();
    }
    for (int i = tid; i < L; i += nt) p.s3[i] = cA[i];
  }

  TL_STAMP();  // 6: after merge
  // ---- fixStep4: bridge unvoiced gaps shorter than 9
  for (int i = tid; i < L; i += nt) cB[i] = cA[i];
  __syncthreads();
  {
    const int nb = tl_boundaries(cA, L, s_blist, 2 * TL_SMAX, p_in.blist, &p.blist, s_scan);
    for (int g = tid; g < nb / 2 - 1; g += nt) {
      const int ed = p.blist[2 * g + 1], st_next = p.blist[2 * (g + 1)];
      const int distance = st_next - ed - 1;
      if (distance >= 9) continue;
      const double tmp0 = cA[ed] + 1;
      const double tmp1 = cA[st_next] - 1;
      const double coefficient = (tmp1 - tmp0) / (distance + 1.0);
      int count = 1;
      for (int j = ed + 1; j <= st_next - 1; ++j) cB[j] = tmp0 + coefficient * count++;
    }
    __syncthreads();
    for (int i = tid; i < L; i += nt) p.s4[i] = cB[i];
  }

  TL_STAMP();  // 7: after fixStep4
  // ---- smoothF0Contour (harvest.cpp:670-703), set-up only: the padded contour and the section table go to
  // global memory; the two zero-phase filter passes run as their own kernels over the whole GPU (this CTA
  // would spend ~40 us of single-SM instruction issue on them)
  const int Lp = L + 2 * TL_LAG;
  double *pad = cA;
  for (int i = tid; i < Lp; i += nt) {
    const double v = (i >= TL_LAG && i < TL_LAG + L) ? cB[i - TL_LAG] : 0.0;
    pad[i] = v;
    p_in.pad[i] = v;
  }
  for (int i = tid; i < L; i += nt) p.out[i] = 0.0;
  __syncthreads();
  const int nbs = tl_boundaries(pad, Lp, s_blist, 2 * TL_SMAX, p_in.blist, &p.blist, s_scan);
  const int nsm = nbs / 2;
  if (nsm > p.maxsec) {
    if (tid == 0) atomicExch(p.error_flag, WB_ERR_UNSUPPORTED);
    return;
  }
  // per section: forward outputs on [st, min(ed + LAG, Lp - 1)], stored at fw[fwoff + (i - st)];
  // work items = TL_CHUNK outputs each.  Table layout (global): st | ed | flen | fw offset | first fwd item |
  // first bwd item, each maxsec + 2 ints.
  if (tid == 0) {
    long long off = 0;
    int items = 0, items_b = 0;
    for (int s = 0; s < nsm; ++s) {
      const int st = p.blist[2 * s], ed = p.blist[2 * s + 1];
      const int flen = min(ed + TL_LAG, Lp - 1) - st + 1;
      p_in.sec_st[s] = st; p_in.sec_ed[s] = ed; p_in.sec_len[s] = flen; p_in.sec_off[s] = (int)off;
      p_in.sec_lo[s] = items;   // first forward work item of this section
      p_in.kb[s] = items_b;     // first backward work item
      off += flen;
      items += (flen + TL_CHUNK - 1) / TL_CHUNK;
      items_b += (ed - st + 1 + TL_CHUNK - 1) / TL_CHUNK;
    }
    p_in.sec_lo[nsm] = items;
    p_in.kb[nsm] = items_b;
    if (off > p.fw_cap) {
      atomicExch(p.error_flag, WB_ERR_UNSUPPORTED);
    } else {
      p_in.smooth_hdr[0] = nsm; p_in.smooth_hdr[1] = items; p_in.smooth_hdr[2] = items_b;
    }
  }
  TL_STAMP();  // 8: smoothing set-up
#undef TL_STAMP
}

// Butterworth coefficients of filteringF0 (harvest.cpp:639-647) and the impulse response of
// 1 / (1 - a0 z^-1 - a1 z^-2) for the look-ahead form of the warm-up
#define TL_B0 0.0078202080334971724
#define TL_B1 0.015640416066994345
#define TL_A0 1.7347257688092754
#define TL_A1 (-0.76600660094326412)

struct SmoothParams {
  const int *hdr;          // nsm, forward items, backward items
  const int *sec_st, *sec_ed, *sec_len, *sec_off, *first_fwd, *first_bwd;
  const double *pad;       // padded contour, L + 2 * TL_LAG
  double *fw;              // forward-pass outputs per section
  double *out;             // smoothed contour, L
};

#define SM_THREADS 128
#define SM_WARPS (SM_THREADS / 32)
#define SM_TABLE 128   // sections whose item table is staged in shared memory
// One WARP per work item (TL_CHUNK outputs of one section plus TL_LAG samples of warm-up): the lanes
// fetch the item's input span with coalesced loads into shared memory, lane 0 then runs the serial
// recurrence out of shared memory (a global load per step would cost more than the arithmetic), and the
// warp stores the outputs.

// section of work item `item`: last s with first[s] <= item
__device__ __forceinline__ int sm_find_section(const int *first, int nsm, int item) {
  int lo = 0, hi = nsm - 1;
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (first[mid] <= item) lo = mid; else hi = mid - 1; }
  return lo;
}

// forward pass (harvest.cpp:649-654); the warm-up before a section start sees the held first value, samples
// after the section end the held last value (the reference's own edge padding)
__global__ void __launch_bounds__(SM_THREADS) smooth_forward_kernel(SmoothParams p) {
  __shared__ double s_x[SM_WARPS][TL_LAG + TL_CHUNK + 1];
  __shared__ int s_first[SM_TABLE];
  const int nsm = p.hdr[0], n_items = p.hdr[1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((int)blockIdx.x * SM_WARPS >= n_items) return;
  const bool staged = nsm <= SM_TABLE;
  if (staged) for (int k = threadIdx.x; k < nsm; k += SM_THREADS) s_first[k] = p.first_fwd[k];
  __syncthreads();
  const int *first = staged ? s_first : p.first_fwd;
  const double b0 = TL_B0, b1 = TL_B1, a0 = TL_A0, a1 = TL_A1;
  const double h1 = a0, h2 = a0 * a0 + a1, h3 = a0 * h2 + a1 * h1, h4 = a0 * h3 + a1 * h2;
  const double *pad = p.pad;
  double *x = s_x[warp];
  for (int item = blockIdx.x * SM_WARPS + warp; item < n_items; item += gridDim.x * SM_WARPS) {
    const int s = sm_find_section(first, nsm, item);
    const int st = p.sec_st[s], ed = p.sec_ed[s], flen = p.sec_len[s];
    const int c = item - first[s];
    const int begin = st + c * TL_CHUNK, end = min(st + flen, begin + TL_CHUNK);
    double *fw = p.fw + p.sec_off[s] - st;
    const int i0 = max(0, begin - TL_LAG);
    const double x_st = pad[st], x_ed = pad[ed];
    __syncwarp();
    for (int k = lane; k < end - i0; k += 32) {
      const int i = i0 + k;
      x[k] = (i < st) ? x_st : (i > ed ? x_ed : pad[i]);
    }
    __syncwarp();
    if (lane == 0) {
      double w0 = 0.0, w1 = 0.0;
      int i = i0;
      // warm-up (only the state matters): four samples per step of the dependency chain
      for (; i + 4 <= begin; i += 4) {
        const double x1 = x[i - i0], x2 = x[i + 1 - i0], x3 = x[i + 2 - i0], x4 = x[i + 3 - i0];
        const double f3 = fma(h2, x1, fma(h1, x2, x3));
        const double f4 = fma(h3, x1, fma(h2, x2, fma(h1, x3, x4)));
        const double n1 = fma(h3, w0, fma(a1 * h2, w1, f3));
        const double n0 = fma(h4, w0, fma(a1 * h3, w1, f4));
        w0 = n0; w1 = n1;
      }
      for (; i < end; ++i) {
        const double wt = x[i - i0] + a0 * w0 + a1 * w1;
        if (i >= begin) x[i - i0] = b0 * wt + b1 * w0 + b0 * w1;   // (the input sample is not needed again)
        w1 = w0; w0 = wt;
      }
    }
    __syncwarp();
    for (int i = begin + lane; i < end; i += 32) fw[i] = x[i - i0];
  }
}

// backward pass (harvest.cpp:656-662): outputs on [st, ed]
__global__ void __launch_bounds__(SM_THREADS) smooth_backward_kernel(SmoothParams p) {
  __shared__ double s_x[SM_WARPS][TL_LAG + TL_CHUNK + 1];
  __shared__ int s_first[SM_TABLE];
  const int nsm = p.hdr[0], n_items = p.hdr[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((int)blockIdx.x * SM_WARPS >= n_items) return;
  const bool staged = nsm <= SM_TABLE;
  if (staged) for (int k = threadIdx.x; k < nsm; k += SM_THREADS) s_first[k] = p.first_bwd[k];
  __syncthreads();
  const int *first = staged ? s_first : p.first_bwd;
  const double b0 = TL_B0, b1 = TL_B1, a0 = TL_A0, a1 = TL_A1;
  const double h1 = a0, h2 = a0 * a0 + a1, h3 = a0 * h2 + a1 * h1, h4 = a0 * h3 + a1 * h2;
  double *x = s_x[warp];
  for (int item = blockIdx.x * SM_WARPS + warp; item < n_items; item += gridDim.x * SM_WARPS) {
    const int s = sm_find_section(first, nsm, item);
    const int st = p.sec_st[s], ed = p.sec_ed[s], flen = p.sec_len[s];
    const int c = item - first[s];
    const int begin = st + c * TL_CHUNK, end = min(ed + 1, begin + TL_CHUNK);  // outputs [begin, end)
    const double *fw = p.fw + p.sec_off[s] - st;
    const int top = min(st + flen - 1, end - 1 + TL_LAG);
    __syncwarp();
    for (int k = lane; k <= top - begin; k += 32) x[k] = fw[begin + k];
    __syncwarp();
    if (lane == 0) {
      double w0 = 0.0, w1 = 0.0;
      int j = top;
      for (; j - 4 >= end - 1; j -= 4) {  // warm-up above the output range, four samples per chain step
        const double x1 = x[j - begin], x2 = x[j - 1 - begin], x3 = x[j - 2 - begin], x4 = x[j - 3 - begin];
        const double f3 = fma(h2, x1, fma(h1, x2, x3));
        const double f4 = fma(h3, x1, fma(h2, x2, fma(h1, x3, x4)));
        const double n1 = fma(h3, w0, fma(a1 * h2, w1, f3));
        const double n0 = fma(h4, w0, fma(a1 * h3, w1, f4));
        w0 = n0; w1 = n1;
      }
      for (; j >= begin; --j) {
        const double wt = x[j - begin] + a0 * w0 + a1 * w1;
        if (j < end) x[j - begin] = b0 * wt + b1 * w0 + b0 * w1;
        w1 = w0; w0 = wt;
      }
    }
    __syncwarp();
    for (int j = begin + lane; j < end; j += 32) p.out[j - TL_LAG] = x[j - begin];
  }
}

// compute(): pick basic_f0 at the frame_period grid (harvest.cpp:199-204)
__global__ void harvest_pick_kernel(const double *__restrict__ basic_f0, int basic_len, double frame_period,
                                    int f0_length, double *__restrict__ tpos, double *__restrict__ f0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= f0_length) return;
  const double t = i * frame_period / 1000.0;
  tpos[i] = t;
  f0[i] = basic_f0[wb_min_i(basic_len - 1, wb_round(t * 1000.0))];
}

}  // namespace

int wb_harvest_tail(WbWorkspace *ws, const double *d_cand, const double *d_score, const int *d_nc, int L,
                    int MC, double *d_f0_out, cudaStream_t stream) {
  TailParams p;
  p.cand = d_cand; p.score = d_score; p.nc = d_nc; p.L = L; p.MC = MC;
  double *c = (double *)ws->get("tl_contours", sizeof(double) * (size_t)L * 5);
  if (!c) return WB_ERR_CUDA;
  p.base = c; p.s1 = c + L; p.s2 = c + 2 * (size_t)L; p.s3 = c + 3 * (size_t)L; p.s4 = c + 4 * (size_t)L;
  p.out = d_f0_out;
  p.blist = (int *)ws->get("tl_blist", sizeof(int) * (L + 2 * TL_LAG + 2));
  p.maxsec = L / 8 + 4;
  int *si = (int *)ws->get("tl_secint", sizeof(int) * (size_t)(p.maxsec + 2) * 9);
  p.sec_sum = (double *)ws->get("tl_secsum", sizeof(double) * (p.maxsec + 2));
  if (!p.blist || !si || !p.sec_sum) return WB_ERR_CUDA;
  const int m = p.maxsec + 2;
  p.sec_st = si; p.sec_ed = si + m; p.sec_lo = si + 2 * m; p.sec_len = si + 3 * m; p.sec_off = si + 4 * m;
  p.kb = si + 5 * m;  // 2 * m
  p.kslot = si + 7 * m; p.order = si + 8 * m;
  p.secbuf_cap = (long long)L + 2LL * (TL_MARGIN + 1) * p.maxsec;
  p.secbuf = (double *)ws->get("tl_secbuf", sizeof(double) * p.secbuf_cap);
  p.pad = (double *)ws->get("tl_pad", sizeof(double) * (L + 2 * TL_LAG));
  p.fw_cap = (long long)L + 2 * TL_LAG + (long long)TL_LAG * p.maxsec;
  p.fw = (double *)ws->get("tl_fw", sizeof(double) * p.fw_cap);
  p.error_flag = ws->error_flag();
  p.clocks = (long long *)ws->get("tl_clocks", sizeof(long long) * 16);
  p.smooth_hdr = (int *)ws->get("tl_smooth_hdr", sizeof(int) * 4);
  p.gA = (double *)ws->get("tl_ga", sizeof(double) * (L + 2 * TL_LAG));
  p.gB = (double *)ws->get("tl_gb", sizeof(double) * (L + 2 * TL_LAG));
  if (!p.gA || !p.gB) return WB_ERR_CUDA;
  if (!p.secbuf || !p.pad || !p.fw || !p.error_flag || !p.clocks || !p.smooth_hdr) return WB_ERR_CUDA;
  WB_LAUNCH("search_base_kernel", search_base_kernel<<<(L * 32 + 255) / 256, 256, 0, stream>>>(d_cand, d_score, d_nc, L, MC, p.base));
  // shared memory: the two contour buffers (+ forward-filter scratch) when they fit (10 s at 1 ms: 2 x 85 KB)
  const int want = 2 * (L + 2 * TL_LAG);
  p.smem_doubles = want <= 25600 ? want : 0;
  const size_t tl_smem_bytes = sizeof(double) * (size_t)p.smem_doubles;
  WB_CUDA_CHECK(cudaFuncSetAttribute(harvest_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl_smem_bytes));
  WB_LAUNCH("harvest_tail_kernel", harvest_tail_kernel<<<1, TL_THREADS, tl_smem_bytes, stream>>>(p));
  WB_CUDA_CHECK(cudaGetLastError());
  {
    SmoothParams q;
    q.hdr = p.smooth_hdr; q.sec_st = p.sec_st; q.sec_ed = p.sec_ed; q.sec_len = p.sec_len; q.sec_off = p.sec_off;
    q.first_fwd = p.sec_lo; q.first_bwd = p.kb; q.pad = p.pad; q.fw = p.fw; q.out = p.out;
    // work items (TL_CHUNK outputs each) are bounded by the padded length plus one partial chunk and one
    // lag of forward outputs per section; the kernels stride over the actual count
    const long long max_items = ((long long)L + 2 * TL_LAG + (long long)TL_LAG * p.maxsec) / TL_CHUNK + p.maxsec + 1;
    const int grid = (int)std::min<long long>((max_items + SM_WARPS - 1) / SM_WARPS, (long long)wb_sm_count() * 8);
    WB_LAUNCH("smooth_forward_kernel", smooth_forward_kernel<<<grid, SM_THREADS, 0, stream>>>(q));
    WB_LAUNCH("smooth_backward_kernel", smooth_backward_kernel<<<grid, SM_THREADS, 0, stream>>>(q));
    WB_CUDA_CHECK(cudaGetLastError());
  }
  return WB_OK;
}

int wb_harvest_pick(const double *d_basic_f0, int basic_len, double frame_period, int f0_length, double *d_tpos,
                    double *d_f0, cudaStream_t stream) {
  WB_LAUNCH("harvest_pick_kernel", harvest_pick_kernel<<<(f0_length + 255) / 256, 256, 0, stream>>>(d_basic_f0, basic_len, frame_period, f0_length,
                                                                  d_tpos, d_f0));
  WB_CUDA_CHECK(cudaGetLastError());
  return WB_OK;
}
