// Source-compatibility header (worldb200) for /root/reference/include/world_fft.hpp.
//
// The reference exposes an FFTW-shaped plan/execute API over Ooura's FFT.  Here a plan records
// the buffers and fft_execute() runs the transform on the GPU through the C-ABI (wb_fft_r2c /
// wb_fft_c2r / wb_fft_c2c): same conventions (forward e^{+i}, backward e^{-i}, unnormalised),
// n a power of two in [128, 16384].  Intended for callers that used these entry points directly;
// the vocoder stages never go through it.
#ifndef WORLD_FFT_HPP
#define WORLD_FFT_HPP

#include "macrodefinitions.hpp"
#include "worldb200.h"

#define FFT_FORWARD 1
#define FFT_BACKWARD 2
#define FFT_ESTIMATE 3

typedef double fft_complex[2];
typedef struct {
  int n;
  int sign;
  unsigned int flags;
  fft_complex *c_in;
  double *in;
  fft_complex *c_out;
  double *out;
  double *input;  // unused (kept for layout compatibility)
  int *ip;        // unused
  double *w;      // unused
} fft_plan;

static inline fft_plan fft_plan_dft_1d(int n, fft_complex *in, fft_complex *out, int sign, unsigned int flags) {
  fft_plan p = {n, sign, flags, in, 0, out, 0, 0, 0, 0};
  return p;
}
static inline fft_plan fft_plan_dft_c2r_1d(int n, fft_complex *in, double *out, unsigned int flags) {
  fft_plan p = {n, FFT_BACKWARD, flags, in, 0, 0, out, 0, 0, 0};
  return p;
}
static inline fft_plan fft_plan_dft_r2c_1d(int n, double *in, fft_complex *out, unsigned int flags) {
  fft_plan p = {n, FFT_FORWARD, flags, 0, in, out, 0, 0, 0, 0};
  return p;
}
static inline void fft_execute(fft_plan p) {
  if (p.sign == FFT_FORWARD) {
    if (p.c_in == 0) wb_fft_r2c(p.in, p.n, 1, &p.c_out[0][0]);
    else wb_fft_c2c(&p.c_in[0][0], p.n, 1, FFT_FORWARD, &p.c_out[0][0]);
  } else {
    if (p.c_out == 0) wb_fft_c2r(&p.c_in[0][0], p.n, 1, p.out);
    else wb_fft_c2c(&p.c_in[0][0], p.n, 1, FFT_BACKWARD, &p.c_out[0][0]);
  }
}
static inline void fft_destroy_plan(fft_plan) {}

#endif
