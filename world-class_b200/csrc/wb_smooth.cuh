// Block-level versions of the reference's spectral helpers, working on shared memory:
//   DCCorrection      /root/reference/src/world_common.cpp:61-80
//   LinearSmoothing   /root/reference/src/world_common.cpp:82-116 (+ :27-52)
//   interp1Q          /root/reference/src/world_matlabfunctions.cpp:220-241
// Every function is called by all threads of the block and ends with __syncthreads().
#pragma once
#include "wb_common.cuh"

// P[0..nc] in shared memory (nc = fft_size / 2).  In-place: P[i] += replica(i).
__device__ inline void wb_dc_correction(double *P, double f0, int fs, int fft_size) {
  const int upper_limit = 2 + static_cast<int>(f0 * fft_size / fs);
  const int n_rep = upper_limit - 1;
  const double x0 = f0;  // f0 - low_frequency_axis[0]
  // (xi - f0) / (-fs / fft_size) is formed as a product with -fft_size / fs: the quotient can differ in the last bit,
  // the interpolant is continuous
  const double inv_dx = -static_cast<double>(fft_size) / fs;
  // Each thread may own several replica points (n_rep can exceed blockDim for huge f0).
  // Read phase and write phase are separated by a barrier because both touch P[0..upper_limit].
  const int nt = blockDim.x;
  double rep[4];
  int cnt = 0;
  for (int i = threadIdx.x; i < n_rep && cnt < 4; i += nt, ++cnt) {
    const double xi = static_cast<double>(i) * fs / fft_size;
    const double qd = (xi - x0) * inv_dx;
    const int base = static_cast<int>(qd);
    const double frac = qd - base;
    // delta_y[x_length - 1] = 0 with x_length = upper_limit + 1 (never reached: base <= upper_limit - 2)
    const double dy = (base >= upper_limit) ? 0.0 : P[base + 1] - P[base];
    rep[cnt] = P[base] + dy * frac;
  }
  __syncthreads();
  cnt = 0;
  for (int i = threadIdx.x; i < n_rep && cnt < 4; i += nt, ++cnt) P[i] += rep[cnt];
  __syncthreads();
}

// in[0..nc] -> out[0..nc] (out may alias in).  seg: scratch of nc + 2*boundary + 1 doubles.
// red: scratch of >= 1024 doubles.  Returns false (uniformly) if seg capacity is exceeded.
__device__ inline bool wb_linear_smoothing(const double *in, double *out, double width, int fs,
                                           int fft_size, double *seg, int seg_capacity, double *red) {
  const int nc = fft_size / 2;
  const int boundary = static_cast<int>(width * fft_size / fs) + 1;
  const int len = nc + boundary * 2 + 1;
  if (len > seg_capacity) return false;
  // fft_size is a power of two: multiplying by its reciprocal is exact, i.e. bit-identical to the division
  const double inv_fft = 1.0 / fft_size;
  // mirroring_spectrum[i] * fs / fft_size  (world_common.cpp:32-47)
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    double v;
    if (i < boundary) v = in[boundary - i];
    else if (i < nc + boundary) v = in[i - boundary];
    else v = in[nc - (i - (nc + boundary))];
    seg[i] = v * fs * inv_fft;
  }
  __syncthreads();
  wb_block_inclusive_scan(seg, len, red);  // mirroring_segment
  const double origin = -(boundary - 0.5) * fs / fft_size;
  // interp1Q (world_matlabfunctions.cpp:220-241) divides by the knot interval fs / fft_size and the result by the
  // width.  The knots are the bin grid shifted by half a bin, so the two positions of bin i are i + c_lo and i + c_hi
  // with constants c: knot index and fraction are formed once (with the reference's expression for bin 0) instead of
  // per bin through a product and a truncation.  The reference's per-bin values equal these up to rounding of the
  // fraction; the interpolant is continuous across a knot, so the results agree to an ulp of the cumulative sum.
  const double inv_interval = static_cast<double>(fft_size) / fs;
  const double inv_width = 1.0 / width;
  const double q_lo = (-width / 2.0 - origin) * inv_interval, q_hi = (-width / 2.0 + width - origin) * inv_interval;
  const int b_lo = static_cast<int>(q_lo), b_hi = static_cast<int>(q_hi);
  const double f_lo = q_lo - b_lo, f_hi = q_hi - b_hi;
  for (int i = threadIdx.x; i <= nc; i += blockDim.x) {
    const double *lo = seg + i + b_lo, *hi = seg + i + b_hi;
    const double low = lo[0] + (lo[1] - lo[0]) * f_lo;
    const double high = hi[0] + (hi[1] - hi[0]) * f_hi;
    out[i] = (high - low) * inv_width;
  }
  __syncthreads();
  return true;
}
