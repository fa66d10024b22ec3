// Shared-memory fp64 FFT for one thread block (sm_100a), replacing the reference's
// Ooura-based world_fft (/root/reference/src/world_fft.cpp:31-167).
//
// Conventions follow the reference wrapper exactly:
//   forward  (r2c, c2c FFT_FORWARD):  X[k] = sum_n x[n] e^{+2 pi i n k / N}   (SIGN = +1)
//   backward (c2r, c2c FFT_BACKWARD): x[n] = sum_k X[k] e^{-2 pi i n k / N}   (SIGN = -1)
//   everything unnormalised (c2r(r2c(x)) = N x).
//
// Layout.  A complex FFT of NC points lives in NC (+ padding) double2 slots of shared
// memory.  Slot padding sidx(i) = i + (i >> 3) makes every pass bank-conflict free for
// 16-byte accesses.  Passes are radix-8 (plus one radix-4/2 clean-up) butterflies held
// in registers; the decimation-in-frequency (DIF) transform reads natural order and
// leaves BIT-REVERSED order, the decimation-in-time (DIT) transform reads bit-reversed
// order and leaves natural order, so forward -> pointwise -> inverse chains need no
// reordering pass.  Twiddles come from a per-size table T[k] = e^{+2 pi i k / NT},
// NT = 2 NC, kept in global memory (L1-resident, read through __ldg).
#pragma once
#include "wb_common.cuh"

__device__ __forceinline__ int wb_sidx(int i) { return i + (i >> 3); }
// number of double2 slots needed for an NC-point complex FFT
__host__ __device__ __forceinline__ int wb_fft_slots(int nc) { return nc + (nc >> 3) + 1; }
// position (in doubles) of real sample j when a real sequence is packed as z[n] = x[2n] + i x[2n+1]
__device__ __forceinline__ int wb_didx(int j) { return 2 * wb_sidx(j >> 1) + (j & 1); }

__device__ __forceinline__ cplx wb_cmul(cplx a, cplx b) {
  cplx r;
  r.x = fma(a.x, b.x, -(a.y * b.y));
  r.y = fma(a.x, b.y, a.y * b.x);
  return r;
}
__device__ __forceinline__ cplx wb_cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx wb_csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx wb_conj(cplx a) { return make_double2(a.x, -a.y); }
// multiply by SIGN * i
template <int SIGN>
__device__ __forceinline__ cplx wb_mul_i(cplx a) {
  return SIGN > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

template <int SIGN>
__device__ __forceinline__ cplx wb_tw(const cplx *__restrict__ T, int idx) {
  cplx w = __ldg(&T[idx]);
  if (SIGN < 0) w.y = -w.y;
  return w;
}

// a[p] <- sum_q a[q] W8^{pq},  W8 = e^{SIGN 2 pi i / 8}
template <int SIGN>
__device__ __forceinline__ void wb_dft8(cplx (&a)[8]) {
  const double r = 0.70710678118654752440;
  cplx b0 = wb_cadd(a[0], a[4]), b4 = wb_csub(a[0], a[4]);
  cplx b1 = wb_cadd(a[1], a[5]), b5 = wb_csub(a[1], a[5]);
  cplx b2 = wb_cadd(a[2], a[6]), b6 = wb_csub(a[2], a[6]);
  cplx b3 = wb_cadd(a[3], a[7]), b7 = wb_csub(a[3], a[7]);
  // b5 *= W8, b6 *= W8^2 = SIGN i, b7 *= W8^3
  b5 = SIGN > 0 ? make_double2((b5.x - b5.y) * r, (b5.x + b5.y) * r)
                : make_double2((b5.x + b5.y) * r, (b5.y - b5.x) * r);
  b6 = wb_mul_i<SIGN>(b6);
  b7 = SIGN > 0 ? make_double2((-b7.x - b7.y) * r, (b7.x - b7.y) * r)
                : make_double2((b7.y - b7.x) * r, (-b7.x - b7.y) * r);
  cplx c0 = wb_cadd(b0, b2), c2 = wb_csub(b0, b2);
  cplx c1 = wb_cadd(b1, b3), c3 = wb_mul_i<SIGN>(wb_csub(b1, b3));
  cplx c4 = wb_cadd(b4, b6), c6 = wb_csub(b4, b6);
  cplx c5 = wb_cadd(b5, b7), c7 = wb_mul_i<SIGN>(wb_csub(b5, b7));
  a[0] = wb_cadd(c0, c1); a[4] = wb_csub(c0, c1);
  a[2] = wb_cadd(c2, c3); a[6] = wb_csub(c2, c3);
  a[1] = wb_cadd(c4, c5); a[5] = wb_csub(c4, c5);
  a[3] = wb_cadd(c6, c7); a[7] = wb_csub(c6, c7);
}

template <int SIGN>
__device__ __forceinline__ void wb_dft4(cplx (&a)[4]) {
  cplx b0 = wb_cadd(a[0], a[2]), b2 = wb_csub(a[0], a[2]);
  cplx b1 = wb_cadd(a[1], a[3]), b3 = wb_mul_i<SIGN>(wb_csub(a[1], a[3]));
  a[0] = wb_cadd(b0, b1); a[2] = wb_csub(b0, b1);
  a[1] = wb_cadd(b2, b3); a[3] = wb_csub(b2, b3);
}

// ---- one DIF pass over sub-blocks of size M (radix R), in place ---------------------
// s: padded slots, N: transform size, T: table with NT entries, tstep = NT / M.
template <int SIGN>
__device__ __forceinline__ void wb_pass_dif8(cplx *s, int N, int M, const cplx *__restrict__ T, int NT) {
  const int m8 = M >> 3, tstep = NT / M;
  for (int u = threadIdx.x; u < (N >> 3); u += blockDim.x) {
    const int j = u & (m8 - 1);
    const int base = ((u - j) << 3) + j;  // (u / m8) * M + j
    cplx a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = s[wb_sidx(base + q * m8)];
    wb_dft8<SIGN>(a);
    if (m8 > 1) {
      const cplx w1 = wb_tw<SIGN>(T, j * tstep);
      const cplx w2 = wb_tw<SIGN>(T, 2 * j * tstep);
      const cplx w4 = wb_tw<SIGN>(T, 4 * j * tstep);
      const cplx w3 = wb_cmul(w1, w2), w5 = wb_cmul(w1, w4), w6 = wb_cmul(w2, w4);
      const cplx w7 = wb_cmul(w3, w4);
      a[1] = wb_cmul(a[1], w1); a[2] = wb_cmul(a[2], w2); a[3] = wb_cmul(a[3], w3);
      a[4] = wb_cmul(a[4], w4); a[5] = wb_cmul(a[5], w5); a[6] = wb_cmul(a[6], w6);
      a[7] = wb_cmul(a[7], w7);
    }
    // slot q' <- A[brev3(q')]
    s[wb_sidx(base + 0 * m8)] = a[0]; s[wb_sidx(base + 1 * m8)] = a[4];
    s[wb_sidx(base + 2 * m8)] = a[2]; s[wb_sidx(base + 3 * m8)] = a[6];
    s[wb_sidx(base + 4 * m8)] = a[1]; s[wb_sidx(base + 5 * m8)] = a[5];
    s[wb_sidx(base + 6 * m8)] = a[3]; s[wb_sidx(base + 7 * m8)] = a[7];
  }
}

template <int SIGN>
__device__ __forceinline__ void wb_pass_dif4(cplx *s, int N, int M, const cplx *__restrict__ T, int NT) {
  const int m4 = M >> 2, tstep = NT / M;
  for (int u = threadIdx.x; u < (N >> 2); u += blockDim.x) {
    const int j = u & (m4 - 1);
    const int base = ((u - j) << 2) + j;
    cplx a[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) a[q] = s[wb_sidx(base + q * m4)];
    wb_dft4<SIGN>(a);
    if (m4 > 1) {
      const cplx w1 = wb_tw<SIGN>(T, j * tstep);
      const cplx w2 = wb_tw<SIGN>(T, 2 * j * tstep);
      const cplx w3 = wb_cmul(w1, w2);
      a[1] = wb_cmul(a[1], w1); a[2] = wb_cmul(a[2], w2); a[3] = wb_cmul(a[3], w3);
    }
    s[wb_sidx(base)] = a[0]; s[wb_sidx(base + m4)] = a[2];
    s[wb_sidx(base + 2 * m4)] = a[1]; s[wb_sidx(base + 3 * m4)] = a[3];
  }
}

template <int SIGN>
__device__ __forceinline__ void wb_pass_dif2(cplx *s, int N, int M, const cplx *__restrict__ T, int NT) {
  const int m2 = M >> 1, tstep = NT / M;
  for (int u = threadIdx.x; u < (N >> 1); u += blockDim.x) {
    const int j = u & (m2 - 1);
    const int base = ((u - j) << 1) + j;
    cplx a0 = s[wb_sidx(base)], a1 = s[wb_sidx(base + m2)];
    cplx d = wb_csub(a0, a1);
    if (m2 > 1) d = wb_cmul(d, wb_tw<SIGN>(T, j * tstep));
    s[wb_sidx(base)] = wb_cadd(a0, a1);
    s[wb_sidx(base + m2)] = d;
  }
}

// ---- DIT passes (transpose of the above) --------------------------------------------
template <int SIGN>
__device__ __forceinline__ void wb_pass_dit8(cplx *s, int N, int M, const cplx *__restrict__ T, int NT) {
  const int m8 = M >> 3, tstep = NT / M;
  for (int u = threadIdx.x; u < (N >> 3); u += blockDim.x) {
    const int j = u & (m8 - 1);
    const int base = ((u - j) << 3) + j;
    cplx a[8];
    // a[r] = slot brev3(r)
    a[0] = s[wb_sidx(base + 0 * m8)]; a[4] = s[wb_sidx(base + 1 * m8)];
    a[2] = s[wb_sidx(base + 2 * m8)]; a[6] = s[wb_sidx(base + 3 * m8)];
    a[1] = s[wb_sidx(base + 4 * m8)]; a[5] = s[wb_sidx(base + 5 * m8)];
    a[3] = s[wb_sidx(base + 6 * m8)]; a[7] = s[wb_sidx(base + 7 * m8)];
    if (m8 > 1) {
      const cplx w1 = wb_tw<SIGN>(T, j * tstep);
      const cplx w2 = wb_tw<SIGN>(T, 2 * j * tstep);
      const cplx w4 = wb_tw<SIGN>(T, 4 * j * tstep);
      const cplx w3 = wb_cmul(w1, w2), w5 = wb_cmul(w1, w4), w6 = wb_cmul(w2, w4);
      const cplx w7 = wb_cmul(w3, w4);
      a[1] = wb_cmul(a[1], w1); a[2] = wb_cmul(a[2], w2); a[3] = wb_cmul(a[3], w3);
      a[4] = wb_cmul(a[4], w4); a[5] = wb_cmul(a[5], w5); a[6] = wb_cmul(a[6], w6);
      a[7] = wb_cmul(a[7], w7);
    }
    wb_dft8<SIGN>(a);
#pragma unroll
    for (int p = 0; p < 8; ++p) s[wb_sidx(base + p * m8)] = a[p];
  }
}

template <int SIGN>
__device__ __forceinline__ void wb_pass_dit4(cplx *s, int N, int M, const cplx *__restrict__ T, int NT) {
  const int m4 = M >> 2, tstep = NT / M;
  for (int u = threadIdx.x; u < (N >> 2); u += blockDim.x) {
    const int j = u & (m4 - 1);
    const int base = ((u - j) << 2) + j;
    cplx a[4];
    a[0] = s[wb_sidx(base)]; a[2] = s[wb_sidx(base + m4)];
    a[1] = s[wb_sidx(base + 2 * m4)]; a[3] = s[wb_sidx(base + 3 * m4)];
    if (m4 > 1) {
      const cplx w1 = wb_tw<SIGN>(T, j * tstep);
      const cplx w2 = wb_tw<SIGN>(T, 2 * j * tstep);
      const cplx w3 = wb_cmul(w1, w2);
      a[1] = wb_cmul(a[1], w1); a[2] = wb_cmul(a[2], w2); a[3] = wb_cmul(a[3], w3);
    }
    wb_dft4<SIGN>(a);
#pragma unroll
    for (int p = 0; p < 4; ++p) s[wb_sidx(base + p * m4)] = a[p];
  }
}

template <int SIGN>
__device__ __forceinline__ void wb_pass_dit2(cplx *s, int N, int M, const cplx *__restrict__ T, int NT) {
  const int m2 = M >> 1, tstep = NT / M;
  for (int u = threadIdx.x; u < (N >> 1); u += blockDim.x) {
    const int j = u & (m2 - 1);
    const int base = ((u - j) << 1) + j;
    cplx a0 = s[wb_sidx(base)], a1 = s[wb_sidx(base + m2)];
    if (m2 > 1) a1 = wb_cmul(a1, wb_tw<SIGN>(T, j * tstep));
    s[wb_sidx(base)] = wb_cadd(a0, a1);
    s[wb_sidx(base + m2)] = wb_csub(a0, a1);
  }
}

// ---- complex transforms -------------------------------------------------------------
// natural order in, bit-reversed order out.  Ends with __syncthreads().
// Caller must __syncthreads() after filling `s`.
template <int SIGN>
__device__ inline void wb_cfft_dif(cplx *s, int N, int log2n, const cplx *__restrict__ T, int NT) {
  const int n8 = log2n / 3, rem = log2n - 3 * n8;
  int M = N;
  for (int p = 0; p < n8; ++p) {
    wb_pass_dif8<SIGN>(s, N, M, T, NT);
    __syncthreads();
    M >>= 3;
  }
  if (rem == 2) { wb_pass_dif4<SIGN>(s, N, 4, T, NT); __syncthreads(); }
  else if (rem == 1) { wb_pass_dif2<SIGN>(s, N, 2, T, NT); __syncthreads(); }
}

// bit-reversed order in, natural order out.  Ends with __syncthreads().
template <int SIGN>
__device__ inline void wb_cfft_dit(cplx *s, int N, int log2n, const cplx *__restrict__ T, int NT) {
  const int n8 = log2n / 3, rem = log2n - 3 * n8;
  int M = 1 << rem;
  if (rem == 2) { wb_pass_dit4<SIGN>(s, N, 4, T, NT); __syncthreads(); }
  else if (rem == 1) { wb_pass_dit2<SIGN>(s, N, 2, T, NT); __syncthreads(); }
  for (int p = 0; p < n8; ++p) {
    M <<= 3;
    wb_pass_dit8<SIGN>(s, N, M, T, NT);
    __syncthreads();
  }
}

__device__ __forceinline__ int wb_brev(int k, int bits) { return (int)(__brev((unsigned)k) >> (32 - bits)); }

// ---- real transforms ------------------------------------------------------------------
// r2c of a real sequence of length N = 2 NC that was packed as z[n] = x[2n] + i x[2n+1]
// into slots wb_sidx(n) (i.e. real sample j at double index wb_didx(j)).
// After the complex DIF transform, `emit(k, X)` is called exactly once for every
// k = 0..NC (X[k] of the length-N real transform, forward sign).  T has NT = N entries.
// The slots are left untouched by the post-processing (emit must not write to s).
template <int SIGN, typename Emit>
__device__ inline void wb_rfft(cplx *s, int NC, int log2nc, const cplx *__restrict__ T, Emit emit) {
  const int N = 2 * NC;
  wb_cfft_dif<SIGN>(s, NC, log2nc, T, N);
  for (int k = threadIdx.x; k <= (NC >> 1); k += blockDim.x) {
    if (k == 0) {
      const cplx z = s[0];
      emit(0, make_double2(z.x + z.y, 0.0));
      emit(NC, make_double2(z.x - z.y, 0.0));
    } else {
      const cplx zk = s[wb_sidx(wb_brev(k, log2nc))];
      const cplx zc = s[wb_sidx(wb_brev(NC - k, log2nc))];
      // E = (Z[k] + conj Z[NC-k]) / 2 ; O = (Z[k] - conj Z[NC-k]) / (2i)
      const cplx E = make_double2(0.5 * (zk.x + zc.x), 0.5 * (zk.y - zc.y));
      const cplx O = make_double2(0.5 * (zk.y + zc.y), -0.5 * (zk.x - zc.x));
      const cplx t = wb_cmul(wb_tw<SIGN>(T, k), O);
      emit(k, wb_cadd(E, t));
      if (k != NC - k) emit(NC - k, wb_conj(wb_csub(E, t)));
    }
  }
  __syncthreads();
}

// c2r (SIGN = -1 for the reference's backward transform): `get(k)` returns X[k] for
// k = 0..NC (Hermitian half; imaginary parts of X[0], X[NC] are ignored like Ooura's
// rdft).  On return real output sample j is at double index wb_didx(j) of `s`.
template <int SIGN, typename Get>
__device__ inline void wb_irfft(cplx *s, int NC, int log2nc, const cplx *__restrict__ T, Get get) {
  const int N = 2 * NC;
  for (int k = threadIdx.x; k <= (NC >> 1); k += blockDim.x) {
    if (k == 0) {
      const double x0 = get(0).x, xn = get(NC).x;
      s[0] = make_double2(x0 + xn, x0 - xn);
    } else {
      const cplx xk = get(k), xc = get(NC - k);
      const cplx A = make_double2(xk.x + xc.x, xk.y - xc.y);               // X[k] + conj X[NC-k]
      const cplx B = wb_cmul(make_double2(xk.x - xc.x, xk.y + xc.y), wb_tw<SIGN>(T, k));
      // Z[k] = A + iB ; Z[NC-k] = conj(A) + i conj(B)
      s[wb_sidx(wb_brev(k, log2nc))] = make_double2(A.x - B.y, A.y + B.x);
      if (k != NC - k) s[wb_sidx(wb_brev(NC - k, log2nc))] = make_double2(A.x + B.y, B.x - A.y);
    }
  }
  __syncthreads();
  wb_cfft_dit<SIGN>(s, NC, log2nc, T, N);
}
