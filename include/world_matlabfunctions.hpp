// Source-compatibility header (worldb200) for the host-side MATLAB-like helpers that callers of
// the reference use between analysis and synthesis (test/test.cpp:228 calls interp1).
// Semantics follow /root/reference/src/world_matlabfunctions.cpp (histc :136-156, interp1
// :158-182, matlab_round :212-214, diff :216-218, fftshift :129-134, interp1Q :220-241); these are
// small host utilities, not part of the GPU hot path.  randn() draws from the library's stream.
#ifndef WORLD_MATLABFUNCTIONS_HPP
#define WORLD_MATLABFUNCTIONS_HPP

#include <vector>

#include "world_common.hpp"
#include "worldb200.h"

inline int matlab_round(double x) { return x > 0 ? static_cast<int>(x + 0.5) : static_cast<int>(x - 0.5); }

inline void fftshift(const double *x, int x_length, double *y) {
  const int half = x_length / 2;
  for (int i = 0; i < half; ++i) {
    y[i] = x[i + half];
    y[i + half] = x[i];
  }
}

inline void diff(const double *x, int x_length, double *y) {
  for (int i = 0; i + 1 < x_length; ++i) y[i] = x[i + 1] - x[i];
}

// index[i] = (1-based) segment of x that holds edges[i]; edges ascending
inline void histc(const double *x, int x_length, const double *edges, int edges_length, int *index) {
  int seg = 1;
  for (int i = 0; i < edges_length; ++i) {
    while (seg < x_length - 1 && !(edges[i] < x[seg])) ++seg;  // first knot above the query, clamped
    index[i] = seg;
  }
}

// piecewise-linear interpolation with linear extrapolation outside the knots
inline void interp1(const double *x, const double *y, int x_length, const double *xi, int xi_length, double *yi) {
  std::vector<int> k(xi_length > 0 ? xi_length : 1);
  histc(x, x_length, xi, xi_length, k.data());
  for (int i = 0; i < xi_length; ++i) {
    const int a = k[i] - 1, b = k[i];
    const double s = (xi[i] - x[a]) / (x[b] - x[a]);
    yi[i] = y[a] + s * (y[b] - y[a]);
  }
}

// interpolation on an equally spaced axis starting at x with step `shift`
inline void interp1Q(double x, double shift, const double *y, int x_length, const double *xi, int xi_length,
                     double *yi) {
  for (int i = 0; i < xi_length; ++i) {
    const int base = static_cast<int>((xi[i] - x) / shift);
    const double fraction = (xi[i] - x) / shift - base;
    const double delta = (base >= x_length - 1) ? 0.0 : y[base + 1] - y[base];
    yi[i] = y[base] + delta * fraction;
  }
}

// zero-phase IIR low-pass + pick of every r-th sample (world_matlabfunctions.cpp:184-210, coefficients :27-125), on
// the GPU with the kernels Harvest uses; r = 2 .. 12; y holds the ceil((x_length + 9 - nbeg) / r) samples the
// reference writes (nbeg = r - r * (x_length / r + 1) + x_length)
inline void decimate(const double *x, int x_length, int r, double *y) { (void)wb_decimate(x, x_length, r, y); }

// next value of the library's randn() stream (same sequence as the reference's generator)
inline double randn(void) {
  double v = 0.0;
  wb_randn_fill(&v, 1);
  return v;
}

#endif
