// Shared-memory fp64 FFT for one thread block (sm_100a), replacing the reference's
// Ooura-based world_fft (/root/reference/src/world_fft.cpp:31-167).
//
// Conventions follow the reference wrapper exactly:
//   forward  (r2c, c2c FFT_FORWARD):  X[k] = sum_n x[n] e^{+2 pi i n k / N}   (SIGN = +1)
//   backward (c2r, c2c FFT_BACKWARD): x[n] = sum_k X[k] e^{-2 pi i n k / N}   (SIGN = -1)
//   everything unnormalised (c2r(r2c(x)) = N x).
//
// Layout.  A complex FFT of N points lives in N (+ padding) double2 slots of shared memory.
// Slot padding sidx(i) = i + (i >> 3) + (i >> 6) + (i >> 9) makes every pass and the
// bit-reversed reads of the real-transform split step bank-conflict free for 16-byte accesses.  Passes are radix-8 (plus one radix-4/2 clean-up) butterflies held in registers;
// the decimation-in-frequency (DIF) transform reads natural order and leaves BIT-REVERSED
// order, the decimation-in-time (DIT) transform reads bit-reversed order and leaves natural
// order, so forward -> pointwise -> inverse chains need no reordering pass.  Twiddles come
// from a table T[k] = e^{+2 pi i k / (2N)} kept in global memory (L1-resident, __ldg).
//
// Everything is templated on log2(N): sub-block sizes, slot offsets (the padded offset of
// element base + q*m is sidx(base) + q*m + (q*m >> 3), a compile-time immediate) and twiddle
// strides fold to constants and every pass is fully unrolled.  Kernels are instantiated per FFT
// size and dispatched on the host.
#pragma once
#include "wb_common.cuh"
#include <type_traits>

// Three-level padding: +1 slot per 8, per 64 and per 512 elements.  The first level keeps the
// small-stride butterfly passes conflict free, the other two spread the bit-reversed access
// pattern of the real-transform split step (lanes hit slots N/32 apart) over all banks.
__device__ __forceinline__ int wb_sidx(int i) { return i + (i >> 3) + (i >> 6) + (i >> 9); }
// number of double2 slots needed for an NC-point complex FFT
__host__ __device__ __forceinline__ constexpr int wb_fft_slots(int nc) { return nc + (nc >> 3) + (nc >> 6) + (nc >> 9) + 1; }
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

// a[p] <- sum_n a[n] W16^{np},  W16 = e^{SIGN 2 pi i / 16}: 4 x 4 decomposition (n = q + 4 t, p = r + 4 s):
// radix-4 over t, twiddle W16^{qr}, radix-4 over q.
template <int SIGN>
__device__ __forceinline__ cplx wb_mul_w16(cplx v, double wr, double wi_abs) {  // v * (wr, SIGN * wi_abs)
  const double wi = SIGN > 0 ? wi_abs : -wi_abs;
  return make_double2(fma(v.x, wr, -(v.y * wi)), fma(v.x, wi, v.y * wr));
}
template <int SIGN>
__device__ __forceinline__ void wb_dft16(cplx (&a)[16]) {
  const double r = 0.70710678118654752440, c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;
  cplx b[4][4];  // b[q][r]
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    cplx t[4] = {a[q], a[q + 4], a[q + 8], a[q + 12]};
    wb_dft4<SIGN>(t);
    b[q][0] = t[0]; b[q][1] = t[1]; b[q][2] = t[2]; b[q][3] = t[3];
  }
  // W16^{qr}
  b[1][1] = wb_mul_w16<SIGN>(b[1][1], c1, s1);                                   // W^1
  b[1][2] = SIGN > 0 ? make_double2((b[1][2].x - b[1][2].y) * r, (b[1][2].x + b[1][2].y) * r)
                     : make_double2((b[1][2].x + b[1][2].y) * r, (b[1][2].y - b[1][2].x) * r);  // W^2
  b[1][3] = wb_mul_w16<SIGN>(b[1][3], s1, c1);                                   // W^3
  b[2][1] = SIGN > 0 ? make_double2((b[2][1].x - b[2][1].y) * r, (b[2][1].x + b[2][1].y) * r)
                     : make_double2((b[2][1].x + b[2][1].y) * r, (b[2][1].y - b[2][1].x) * r);  // W^2
  b[2][2] = wb_mul_i<SIGN>(b[2][2]);                                             // W^4
  b[2][3] = SIGN > 0 ? make_double2((-b[2][3].x - b[2][3].y) * r, (b[2][3].x - b[2][3].y) * r)
                     : make_double2((b[2][3].y - b[2][3].x) * r, (-b[2][3].x - b[2][3].y) * r);  // W^6
  b[3][1] = wb_mul_w16<SIGN>(b[3][1], s1, c1);                                   // W^3
  b[3][2] = SIGN > 0 ? make_double2((-b[3][2].x - b[3][2].y) * r, (b[3][2].x - b[3][2].y) * r)
                     : make_double2((b[3][2].y - b[3][2].x) * r, (-b[3][2].x - b[3][2].y) * r);  // W^6
  b[3][3] = wb_mul_w16<SIGN>(b[3][3], -c1, -s1);                                 // W^9 = -W^1
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    cplx t[4] = {b[0][rr], b[1][rr], b[2][rr], b[3][rr]};
    wb_dft4<SIGN>(t);
    a[rr] = t[0]; a[rr + 4] = t[1]; a[rr + 8] = t[2]; a[rr + 12] = t[3];
  }
}

// padded offset of element (base + q * m) relative to sidx(base): the three shift terms separate
// because base = M * block + j with j < m = M / 8, so (base mod 2^s) + (q m mod 2^s) never carries
// for s = 3, 6, 9 (for M <= 2^s the whole sub-block lies inside one 2^s-aligned group).
#define WB_OFF(q, m) ((q) * (m) + (((q) * (m)) >> 3) + (((q) * (m)) >> 6) + (((q) * (m)) >> 9))

// twiddles w^1..w^7 of one radix-8 butterfly from three table loads
template <int SIGN>
__device__ __forceinline__ void wb_twiddles8(const cplx *__restrict__ T, int t1, cplx (&w)[8]) {
  w[1] = wb_tw<SIGN>(T, t1);
  w[2] = wb_tw<SIGN>(T, 2 * t1);
  w[4] = wb_tw<SIGN>(T, 4 * t1);
  w[3] = wb_cmul(w[1], w[2]);
  w[5] = wb_cmul(w[1], w[4]);
  w[6] = wb_cmul(w[2], w[4]);
  w[7] = wb_cmul(w[3], w[4]);
}

// `in(i, q)` of a DIF pass supplies element i of the input (q = its position inside the butterfly, a
// compile-time constant after unrolling, so the functor may index per-thread register arrays with it): the
// default reads the slots, a caller-provided functor fuses the producer of the data into the first pass
// (no store + reload of the input, and zero padding costs nothing).
struct WbFromSlots {};

// ---- one DIF pass over sub-blocks of size M (radix 8), in place; table has 2N entries -------
// OWN16: clean-up pass (M = 8) of the radix-16 plan in WL mode: thread t takes the two butterflies inside its 16 points
template <int SIGN, int N, int M, typename In = WbFromSlots, bool OWN16 = false>
__device__ __forceinline__ void wb_pass_dif8(cplx *s, const cplx *__restrict__ T, In in = In()) {
  constexpr int m8 = M / 8, tstep = 2 * N / M;
  static_assert(!OWN16 || M == 8, "ownership mapping is for the twiddle-free clean-up pass");
  const int u_begin = OWN16 ? 2 * (int)threadIdx.x : (int)threadIdx.x;
  const int u_step = OWN16 ? 1 : (int)blockDim.x;
  const int u_end = OWN16 ? (((int)threadIdx.x < N / 16) ? u_begin + 2 : u_begin) : N / 8;
  for (int u = u_begin; u < u_end; u += u_step) {
    const int j = u & (m8 - 1);
    const int base = ((u - j) << 3) + j;  // (u / m8) * M + j
    cplx *sp = s + wb_sidx(base);
    cplx a[8];
    if constexpr (std::is_same<In, WbFromSlots>::value) {
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = sp[WB_OFF(q, m8)];
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = in(base + q * m8, q);
    }
    wb_dft8<SIGN>(a);
    if (m8 > 1) {
      cplx w[8];
      wb_twiddles8<SIGN>(T, j * tstep, w);
#pragma unroll
      for (int q = 1; q < 8; ++q) a[q] = wb_cmul(a[q], w[q]);
    }
    // slot q' <- A[brev3(q')]
    sp[WB_OFF(0, m8)] = a[0]; sp[WB_OFF(1, m8)] = a[4]; sp[WB_OFF(2, m8)] = a[2]; sp[WB_OFF(3, m8)] = a[6];
    sp[WB_OFF(4, m8)] = a[1]; sp[WB_OFF(5, m8)] = a[5]; sp[WB_OFF(6, m8)] = a[3]; sp[WB_OFF(7, m8)] = a[7];
  }
}

// a[p] *= w^p, p = 1..15, with w^1, w^2, w^4, w^8 from the table and the rest by products
template <int SIGN>
__device__ __forceinline__ void wb_apply_twiddles16(const cplx *__restrict__ T, int t1, cplx (&a)[16]) {
  const cplx w1 = wb_tw<SIGN>(T, t1), w2 = wb_tw<SIGN>(T, 2 * t1), w4 = wb_tw<SIGN>(T, 4 * t1), w8 = wb_tw<SIGN>(T, 8 * t1);
  a[1] = wb_cmul(a[1], w1);
  a[2] = wb_cmul(a[2], w2);
  const cplx w3 = wb_cmul(w1, w2);
  a[3] = wb_cmul(a[3], w3);
  a[4] = wb_cmul(a[4], w4);
  const cplx w5 = wb_cmul(w1, w4);
  a[5] = wb_cmul(a[5], w5);
  const cplx w6 = wb_cmul(w2, w4);
  a[6] = wb_cmul(a[6], w6);
  const cplx w7 = wb_cmul(w3, w4);
  a[7] = wb_cmul(a[7], w7);
  a[8] = wb_cmul(a[8], w8);
  a[9] = wb_cmul(a[9], wb_cmul(w1, w8));
  a[10] = wb_cmul(a[10], wb_cmul(w2, w8));
  a[11] = wb_cmul(a[11], wb_cmul(w3, w8));
  a[12] = wb_cmul(a[12], wb_cmul(w4, w8));
  a[13] = wb_cmul(a[13], wb_cmul(w5, w8));
  a[14] = wb_cmul(a[14], wb_cmul(w6, w8));
  a[15] = wb_cmul(a[15], wb_cmul(w7, w8));
}

__device__ __forceinline__ constexpr int wb_brev4c(int q) { return ((q & 1) << 3) | ((q & 2) << 1) | ((q & 4) >> 1) | ((q & 8) >> 3); }

// ---- radix-16 DIF pass over sub-blocks of size M, in place (input functor: see WbFromSlots)
template <int SIGN, int N, int M, typename In>
__device__ __forceinline__ void wb_pass_dif16(cplx *s, const cplx *__restrict__ T, In in) {
  constexpr int m = M / 16, tstep = 2 * N / M;
  for (int u = threadIdx.x; u < N / 16; u += blockDim.x) {
    const int j = u & (m - 1);
    const int base = ((u - j) << 4) + j;  // (u / m) * M + j
    cplx *sp = s + wb_sidx(base);
    cplx a[16];
    if constexpr (std::is_same<In, WbFromSlots>::value) {
#pragma unroll
      for (int q = 0; q < 16; ++q) a[q] = sp[WB_OFF(q, m)];
    } else {
#pragma unroll
      for (int q = 0; q < 16; ++q) a[q] = in(base + q * m, q);
    }
    wb_dft16<SIGN>(a);
    if (m > 1) wb_apply_twiddles16<SIGN>(T, j * tstep, a);
    // slot q' <- A[brev4(q')]
#pragma unroll
    for (int q = 0; q < 16; ++q) sp[WB_OFF(q, m)] = a[wb_brev4c(q)];
  }
}

// ---- radix-16 DIT pass (transpose of the above)
template <int SIGN, int N, int M>
__device__ __forceinline__ void wb_pass_dit16(cplx *s, const cplx *__restrict__ T) {
  constexpr int m = M / 16, tstep = 2 * N / M;
  for (int u = threadIdx.x; u < N / 16; u += blockDim.x) {
    const int j = u & (m - 1);
    const int base = ((u - j) << 4) + j;
    cplx *sp = s + wb_sidx(base);
    cplx a[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) a[wb_brev4c(q)] = sp[WB_OFF(q, m)];
    if (m > 1) wb_apply_twiddles16<SIGN>(T, j * tstep, a);
    wb_dft16<SIGN>(a);
#pragma unroll
    for (int p = 0; p < 16; ++p) sp[WB_OFF(p, m)] = a[p];
  }
}

// Warp-local late passes (template flag WL of the plans below).  With ONE radix-R butterfly per thread and pass
// (N / R <= blockDim, butterfly u on thread u), the butterflies of a sub-block of M <= 32 R points sit on M / R
// consecutive threads of one warp, and so do all butterflies of the later (DIF) / earlier (DIT) passes inside that
// sub-block: between those passes __syncwarp() orders the shared-memory traffic and the block-wide barrier is
// only needed around the passes that span warps.  The radix-8 / 4 / 2 clean-up pass has more butterflies than
// threads; in WL mode thread t takes the ones inside its own R consecutive points [R t, R (t + 1)).
// OWN = points per thread = R of the plan; RL = radix of the clean-up pass.
template <int OWN, int RL>
__device__ __forceinline__ int wb_wl_first(int k) { return (int)threadIdx.x * (OWN / RL) + k; }

// final radix-4 / radix-2 pass (sub-block size 4 or 2: no twiddles)
template <int SIGN, int N, int OWN = 0>
__device__ __forceinline__ void wb_pass_dif4_last(cplx *s) {
  if constexpr (OWN > 0) {
    if ((int)threadIdx.x < N / OWN) {
#pragma unroll
      for (int k = 0; k < OWN / 4; ++k) {
        cplx *sp = s + wb_sidx(4 * wb_wl_first<OWN, 4>(k));
        cplx a[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] = sp[q];
        wb_dft4<SIGN>(a);
        sp[0] = a[0]; sp[1] = a[2]; sp[2] = a[1]; sp[3] = a[3];
      }
    }
  } else {
  for (int u = threadIdx.x; u < N / 4; u += blockDim.x) {
    cplx *sp = s + wb_sidx(4 * u);
    cplx a[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) a[q] = sp[q];
    wb_dft4<SIGN>(a);
    sp[0] = a[0]; sp[1] = a[2]; sp[2] = a[1]; sp[3] = a[3];
  }
  }
}
template <int SIGN, int N, int OWN = 0>
__device__ __forceinline__ void wb_pass_dif2_last(cplx *s) {
  if constexpr (OWN > 0) {
    if ((int)threadIdx.x < N / OWN) {
#pragma unroll
      for (int k = 0; k < OWN / 2; ++k) {
        cplx *sp = s + wb_sidx(2 * wb_wl_first<OWN, 2>(k));
        const cplx a0 = sp[0], a1 = sp[1];
        sp[0] = wb_cadd(a0, a1);
        sp[1] = wb_csub(a0, a1);
      }
    }
  } else {
  for (int u = threadIdx.x; u < N / 2; u += blockDim.x) {
    cplx *sp = s + wb_sidx(2 * u);
    const cplx a0 = sp[0], a1 = sp[1];
    sp[0] = wb_cadd(a0, a1);
    sp[1] = wb_csub(a0, a1);
  }
  }
}

// ---- DIT passes (transpose of the above) --------------------------------------------
template <int SIGN, int N, int M, bool OWN16 = false>
__device__ __forceinline__ void wb_pass_dit8(cplx *s, const cplx *__restrict__ T) {
  constexpr int m8 = M / 8, tstep = 2 * N / M;
  static_assert(!OWN16 || M == 8, "ownership mapping is for the twiddle-free first pass");
  const int u_begin = OWN16 ? 2 * (int)threadIdx.x : (int)threadIdx.x;
  const int u_step = OWN16 ? 1 : (int)blockDim.x;
  const int u_end = OWN16 ? (((int)threadIdx.x < N / 16) ? u_begin + 2 : u_begin) : N / 8;
  for (int u = u_begin; u < u_end; u += u_step) {
    const int j = u & (m8 - 1);
    const int base = ((u - j) << 3) + j;
    cplx *sp = s + wb_sidx(base);
    cplx a[8];
    // a[r] = slot brev3(r)
    a[0] = sp[WB_OFF(0, m8)]; a[4] = sp[WB_OFF(1, m8)]; a[2] = sp[WB_OFF(2, m8)]; a[6] = sp[WB_OFF(3, m8)];
    a[1] = sp[WB_OFF(4, m8)]; a[5] = sp[WB_OFF(5, m8)]; a[3] = sp[WB_OFF(6, m8)]; a[7] = sp[WB_OFF(7, m8)];
    if (m8 > 1) {
      cplx w[8];
      wb_twiddles8<SIGN>(T, j * tstep, w);
#pragma unroll
      for (int q = 1; q < 8; ++q) a[q] = wb_cmul(a[q], w[q]);
    }
    wb_dft8<SIGN>(a);
#pragma unroll
    for (int p = 0; p < 8; ++p) sp[WB_OFF(p, m8)] = a[p];
  }
}
template <int SIGN, int N, int OWN = 0>
__device__ __forceinline__ void wb_pass_dit4_first(cplx *s) {
  if constexpr (OWN > 0) {
    if ((int)threadIdx.x < N / OWN) {
#pragma unroll
      for (int k = 0; k < OWN / 4; ++k) {
        cplx *sp = s + wb_sidx(4 * wb_wl_first<OWN, 4>(k));
        cplx a[4];
        a[0] = sp[0]; a[2] = sp[1]; a[1] = sp[2]; a[3] = sp[3];
        wb_dft4<SIGN>(a);
#pragma unroll
        for (int p = 0; p < 4; ++p) sp[p] = a[p];
      }
    }
  } else {
  for (int u = threadIdx.x; u < N / 4; u += blockDim.x) {
    cplx *sp = s + wb_sidx(4 * u);
    cplx a[4];
    a[0] = sp[0]; a[2] = sp[1]; a[1] = sp[2]; a[3] = sp[3];
    wb_dft4<SIGN>(a);
#pragma unroll
    for (int p = 0; p < 4; ++p) sp[p] = a[p];
  }
  }
}
template <int SIGN, int N, int OWN = 0>
__device__ __forceinline__ void wb_pass_dit2_first(cplx *s) {
  if constexpr (OWN > 0) {
    if ((int)threadIdx.x < N / OWN) {
#pragma unroll
      for (int k = 0; k < OWN / 2; ++k) {
        cplx *sp = s + wb_sidx(2 * wb_wl_first<OWN, 2>(k));
        const cplx a0 = sp[0], a1 = sp[1];
        sp[0] = wb_cadd(a0, a1);
        sp[1] = wb_csub(a0, a1);
      }
    }
  } else {
  for (int u = threadIdx.x; u < N / 2; u += blockDim.x) {
    cplx *sp = s + wb_sidx(2 * u);
    const cplx a0 = sp[0], a1 = sp[1];
    sp[0] = wb_cadd(a0, a1);
    sp[1] = wb_csub(a0, a1);
  }
  }
}

// Pass plans (template parameter R):
//   R = 16: radix 16 while the sub-block size allows it, then ONE twiddle-free clean-up pass of radix
//           8 / 4 / 2 (4096 = 16.16.16, 2048 = 16.16.8, 1024 = 16.16.4, ...): three passes over shared
//           memory where radix 8 needs four, at ~100-128 registers per thread (two 256-thread CTAs per SM).
//   R = 8:  radix 8 plus a radix-4 / 2 clean-up pass, ~60-75 registers: for kernels whose shared-memory
//           footprint allows three or more CTAs per SM, where occupancy is worth more than the saved pass.
#ifndef WB_FFT_DEFAULT_RADIX
#define WB_FFT_DEFAULT_RADIX 8
#endif
// sync between the pass that has just produced sub-blocks of M / R points... (WL: see above)
template <bool WARP>
__device__ __forceinline__ void wb_fft_sync() {
  if constexpr (WARP) __syncwarp(); else __syncthreads();
}

template <int SIGN, int N, int M, int R, bool WL = false>
struct WbDifPasses {
  template <typename In>
  static __device__ __forceinline__ void run(cplx *s, const cplx *__restrict__ T, In in) {
    // after a pass over sub-blocks of M points the next pass works inside sub-blocks of M / radix points: warp-local
    // when the M-point sub-block already sits on one warp (M <= 32 R) -- except after the LAST pass (callers
    // read across the whole transform)
    if constexpr (R == 16 && M >= 16) {
      wb_pass_dif16<SIGN, N, M>(s, T, in);
      wb_fft_sync<(WL && M <= 32 * 16 && M > 16)>();
      WbDifPasses<SIGN, N, M / 16, R, WL>::run(s, T, WbFromSlots());
    } else {
      static_assert(M >= 8 || std::is_same<In, WbFromSlots>::value, "fused input needs N >= 8");
      if constexpr (M >= 8) {
        constexpr bool own16 = WL && R == 16 && M == 8;        // clean-up pass of the radix-16 plan
        wb_pass_dif8<SIGN, N, M, In, own16>(s, T, in);
        wb_fft_sync<(WL && R == 8 && M <= 32 * 8 && M > 8)>();
        WbDifPasses<SIGN, N, M / 8, R, WL>::run(s, T, WbFromSlots());
      } else if constexpr (M == 4) {
        wb_pass_dif4_last<SIGN, N, (WL ? R : 0)>(s);
        __syncthreads();
      } else if constexpr (M == 2) {
        wb_pass_dif2_last<SIGN, N, (WL ? R : 0)>(s);
        __syncthreads();
      }
    }
  }
};

template <int SIGN, int N, int M, int R, bool WL = false>  // M = sub-block size produced by the passes done so far
struct WbDitPasses {
  static __device__ __forceinline__ void run(cplx *s, const cplx *__restrict__ T) {
    if constexpr (M < N) {
      // the pass below reads sub-blocks of M points and produces sub-blocks of M R points; the pass after it is
      // warp-local with it when those M R points sit on one warp and it is not the last pass
      if constexpr (R == 16) {
        wb_pass_dit16<SIGN, N, M * 16>(s, T);
        wb_fft_sync<(WL && M * 16 * 16 <= 32 * 16 && M * 16 < N)>();
        WbDitPasses<SIGN, N, M * 16, R, WL>::run(s, T);
      } else {
        wb_pass_dit8<SIGN, N, M * 8>(s, T);
        wb_fft_sync<(WL && M * 8 * 8 <= 32 * 8 && M * 8 < N)>();
        WbDitPasses<SIGN, N, M * 8, R, WL>::run(s, T);
      }
    }
  }
};

// ---- complex transforms: N = 2^LOG2N points, table T with 2N entries -------------------------
// natural order in, bit-reversed order out.  Ends with __syncthreads().
// Caller must __syncthreads() after filling `s`.
// WL (all transforms below): warp-local late passes, see above; the caller guarantees N / R <= blockDim.
template <int SIGN, int LOG2N, int R = WB_FFT_DEFAULT_RADIX, bool WL = false>
__device__ __forceinline__ void wb_cfft_dif_t(cplx *s, const cplx *__restrict__ T) {
  WbDifPasses<SIGN, (1 << LOG2N), (1 << LOG2N), R, WL>::run(s, T, WbFromSlots());
}

// Radix-16 plan with element i of the input coming from in(i, q) instead of the slots (LOG2N >= 4).  The slots
// are only written: the caller must make sure nobody still reads them (a __syncthreads() after their last use).
template <int SIGN, int LOG2N, int R = 16, bool WL = false, typename In>
__device__ __forceinline__ void wb_cfft_dif_in_t(cplx *s, const cplx *__restrict__ T, In in) {
  WbDifPasses<SIGN, (1 << LOG2N), (1 << LOG2N), R, WL>::run(s, T, in);
}

// bit-reversed order in, natural order out.  Ends with __syncthreads().
template <int SIGN, int LOG2N, int R = WB_FFT_DEFAULT_RADIX, bool WL = false>
__device__ __forceinline__ void wb_cfft_dit_t(cplx *s, const cplx *__restrict__ T) {
  constexpr int N = 1 << LOG2N;
  // (the pass after the clean-up pass works inside sub-blocks of at most 128 points: always on one warp)
  if constexpr (R == 16) {
    constexpr int rem = LOG2N % 4;
    if constexpr (rem == 3) { wb_pass_dit8<SIGN, N, 8, WL>(s, T); wb_fft_sync<(WL && N > 8)>(); }
    if constexpr (rem == 2) { wb_pass_dit4_first<SIGN, N, (WL ? 16 : 0)>(s); wb_fft_sync<(WL && N > 4)>(); }
    if constexpr (rem == 1) { wb_pass_dit2_first<SIGN, N, (WL ? 16 : 0)>(s); wb_fft_sync<(WL && N > 2)>(); }
    WbDitPasses<SIGN, N, (1 << rem), R, WL>::run(s, T);
  } else {
    constexpr int rem = LOG2N % 3;
    if constexpr (rem == 2) { wb_pass_dit4_first<SIGN, N, (WL ? 8 : 0)>(s); wb_fft_sync<(WL && N > 4)>(); }
    if constexpr (rem == 1) { wb_pass_dit2_first<SIGN, N, (WL ? 8 : 0)>(s); wb_fft_sync<(WL && N > 2)>(); }
    WbDitPasses<SIGN, N, (1 << rem), R, WL>::run(s, T);
  }
}

__device__ __forceinline__ int wb_brev(int k, int bits) { return (int)(__brev((unsigned)k) >> (32 - bits)); }

// ---- real transforms ------------------------------------------------------------------
// r2c of a real sequence of length 2 NC (NC = 2^LOG2NC) that was packed as
// z[n] = x[2n] + i x[2n+1] into slots wb_sidx(n) (real sample j at double index wb_didx(j)).
// After the complex DIF transform, `emit(k, X)` is called exactly once for every k = 0..NC
// (X[k] of the real transform, forward sign).  T has 2 NC entries.
// The slots are left untouched by the post-processing (emit must not write to s).
template <int SIGN, int LOG2NC, int R = WB_FFT_DEFAULT_RADIX, bool WL = false, typename Emit>
__device__ __forceinline__ void wb_rfft_t(cplx *s, const cplx *__restrict__ T, Emit emit) {
  constexpr int NC = 1 << LOG2NC;
  wb_cfft_dif_t<SIGN, LOG2NC, R, WL>(s, T);
  for (int k = threadIdx.x; k <= (NC >> 1); k += blockDim.x) {
    if (k == 0) {
      const cplx z = s[0];
      emit(0, make_double2(z.x + z.y, 0.0));
      emit(NC, make_double2(z.x - z.y, 0.0));
    } else {
      const cplx zk = s[wb_sidx(wb_brev(k, LOG2NC))];
      const cplx zc = s[wb_sidx(wb_brev(NC - k, LOG2NC))];
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
template <int SIGN, int LOG2NC, int R = WB_FFT_DEFAULT_RADIX, bool WL = false, typename Get>
__device__ __forceinline__ void wb_irfft_t(cplx *s, const cplx *__restrict__ T, Get get) {
  constexpr int NC = 1 << LOG2NC;
  for (int k = threadIdx.x; k <= (NC >> 1); k += blockDim.x) {
    if (k == 0) {
      const double x0 = get(0).x, xn = get(NC).x;
      s[0] = make_double2(x0 + xn, x0 - xn);
    } else {
      const cplx xk = get(k), xc = get(NC - k);
      const cplx A = make_double2(xk.x + xc.x, xk.y - xc.y);               // X[k] + conj X[NC-k]
      const cplx B = wb_cmul(make_double2(xk.x - xc.x, xk.y + xc.y), wb_tw<SIGN>(T, k));
      // Z[k] = A + iB ; Z[NC-k] = conj(A) + i conj(B)
      s[wb_sidx(wb_brev(k, LOG2NC))] = make_double2(A.x - B.y, A.y + B.x);
      if (k != NC - k) s[wb_sidx(wb_brev(NC - k, LOG2NC))] = make_double2(A.x + B.y, B.x - A.y);
    }
  }
  __syncthreads();
  wb_cfft_dit_t<SIGN, LOG2NC, R, WL>(s, T);
}

// ---- host-side dispatch helper: call F.template operator()<LOG2N>() for a runtime log2n ------
#define WB_DISPATCH_LOG2(LOG2, LO, HI, ...)                              \
  [&]() -> int {                                                        \
    switch (LOG2) {                                                     \
      case 6: if (6 >= LO && 6 <= HI) { constexpr int L2 = (6 >= LO && 6 <= HI) ? 6 : LO; __VA_ARGS__; return WB_OK; } break;   \
      case 7: if (7 >= LO && 7 <= HI) { constexpr int L2 = (7 >= LO && 7 <= HI) ? 7 : LO; __VA_ARGS__; return WB_OK; } break;   \
      case 8: if (8 >= LO && 8 <= HI) { constexpr int L2 = (8 >= LO && 8 <= HI) ? 8 : LO; __VA_ARGS__; return WB_OK; } break;   \
      case 9: if (9 >= LO && 9 <= HI) { constexpr int L2 = (9 >= LO && 9 <= HI) ? 9 : LO; __VA_ARGS__; return WB_OK; } break;   \
      case 10: if (10 >= LO && 10 <= HI) { constexpr int L2 = (10 >= LO && 10 <= HI) ? 10 : LO; __VA_ARGS__; return WB_OK; } break; \
      case 11: if (11 >= LO && 11 <= HI) { constexpr int L2 = (11 >= LO && 11 <= HI) ? 11 : LO; __VA_ARGS__; return WB_OK; } break; \
      case 12: if (12 >= LO && 12 <= HI) { constexpr int L2 = (12 >= LO && 12 <= HI) ? 12 : LO; __VA_ARGS__; return WB_OK; } break; \
      case 13: if (13 >= LO && 13 <= HI) { constexpr int L2 = (13 >= LO && 13 <= HI) ? 13 : LO; __VA_ARGS__; return WB_OK; } break; \
      case 14: if (14 >= LO && 14 <= HI) { constexpr int L2 = (14 >= LO && 14 <= HI) ? 14 : LO; __VA_ARGS__; return WB_OK; } break; \
      default: break;                                                   \
    }                                                                   \
    return WB_ERR_UNSUPPORTED;                                          \
  }()
