// Source-compatibility header (worldb200) for the parts of /root/reference/include/world_common.hpp
// (:17-125): the min/max helpers, GetSuitableFFTSize, GetSafeAperiodicity, the spectral helpers DCCorrection /
// LinearSmoothing / NuttallWindow and the FFT scratch structs (ForwardRealFFT ... MinimumPhaseAnalysis).
// The vocoder stages of this library never go through these (their transforms and smoothing passes live in GPU
// shared memory); they are here for callers that used the reference's public helpers directly.  The helpers are
// small host functions with the reference's semantics (world_common.cpp:27-126); the structs keep the
// reference's fields and run their transforms through world_fft.hpp, i.e. on the GPU.
#ifndef WORLD_COMMON_HPP
#define WORLD_COMMON_HPP

#include <cmath>
#include <vector>

#include "macrodefinitions.hpp"
#include "world_constantnumbers.hpp"
#include "world_fft.hpp"

inline int MyMaxInt(int x, int y) { return x > y ? x : y; }
inline double MyMaxDouble(double x, double y) { return x > y ? x : y; }
inline int MyMinInt(int x, int y) { return x < y ? x : y; }
inline double MyMinDouble(double x, double y) { return x < y ? x : y; }

// world_common.cpp:56-59
inline int GetSuitableFFTSize(int sample) {
  return static_cast<int>(std::pow(2.0, static_cast<int>(std::log(static_cast<double>(sample)) / world::kLog2) + 1.0));
}

// world_common.hpp:123-125
inline double GetSafeAperiodicity(double x) { return MyMaxDouble(0.001, MyMinDouble(0.999999999999, x)); }

// value at `at` of the polyline through (x0 + k * step, y[k]), k = 0 .. n-1; the last knot repeats beyond the end
// (what the reference's interp1Q, world_matlabfunctions.cpp:220-241, computes per query)
inline double wb_polyline_at(double x0, double step, const double *y, int n, double at) {
  const double q = (at - x0) / step;
  const int k = static_cast<int>(q);
  const double rise = (k >= n - 1) ? 0.0 : y[k + 1] - y[k];
  return y[k] + rise * (q - k);
}

// world_common.cpp:61-80: the power below f0 receives the replica mirrored about f0
inline void DCCorrection(const double *input, double current_f0, int fs, int fft_size, double *output) {
  const int upper_limit = 2 + static_cast<int>(current_f0 * fft_size / fs);
  const double step = -static_cast<double>(fs) / fft_size;
  for (int i = 0; i < upper_limit - 1; ++i) {
    const double f = static_cast<double>(i) * fs / fft_size;
    output[i] = input[i] + wb_polyline_at(current_f0, step, input, upper_limit + 1, f);
  }
}

// world_common.cpp:82-116 (+ :27-52): rectangular smoothing of `width` Hz through the running integral of the
// spectrum mirrored at both ends
inline void LinearSmoothing(const double *input, double width, int fs, int fft_size, double *output) {
  const int half = fft_size / 2;
  const int boundary = static_cast<int>(width * fft_size / fs) + 1;
  const int n = half + 2 * boundary + 1;
  std::vector<double> integral(n);
  double run = 0.0;
  for (int i = 0; i < n; ++i) {
    const int k = i - boundary;                               // spectrum bin under slot i, reflected at 0 and half
    const double v = input[k < 0 ? -k : (k > half ? 2 * half - k : k)];
    run = (i == 0) ? v * fs / fft_size : v * fs / fft_size + run;
    integral[i] = run;
  }
  const double origin = -(boundary - 0.5) * fs / fft_size;
  const double step = static_cast<double>(fs) / fft_size;
  for (int i = 0; i <= half; ++i) {
    double f = static_cast<double>(i) / fft_size * fs - width / 2.0;
    const double low = wb_polyline_at(origin, step, integral.data(), n, f);
    f += width;
    const double high = wb_polyline_at(origin, step, integral.data(), n, f);
    output[i] = (high - low) / width;
  }
}

// world_common.cpp:118-126
inline void NuttallWindow(int y_length, double *y) {
  for (int i = 0; i < y_length; ++i) {
    const double t = i / (y_length - 1.0);
    y[i] = 0.355768 - 0.487396 * std::cos(2.0 * world::kPi * t) + 0.144232 * std::cos(4.0 * world::kPi * t) -
           0.012604 * std::cos(6.0 * world::kPi * t);
  }
}

// ---- FFT scratch structs (world_common.hpp:17-61, world_common.cpp:131-233): same fields; the plans run on the GPU
typedef struct ForwardRealFFT {
  int fft_size;
  double *waveform;
  fft_complex *spectrum;
  fft_plan forward_fft;
  void initialize(int n) {
    fft_size = n;
    waveform = new double[n];
    spectrum = new fft_complex[n];
    forward_fft = fft_plan_dft_r2c_1d(n, waveform, spectrum, FFT_ESTIMATE);
  }
  void destroy() { fft_destroy_plan(forward_fft); delete[] spectrum; delete[] waveform; }
} ForwardRealFFT;

typedef struct InverseRealFFT {
  int fft_size;
  double *waveform;
  fft_complex *spectrum;
  fft_plan inverse_fft;
  void initialize(int n) {
    fft_size = n;
    waveform = new double[n];
    spectrum = new fft_complex[n];
    inverse_fft = fft_plan_dft_c2r_1d(n, spectrum, waveform, FFT_ESTIMATE);
  }
  void destroy() { fft_destroy_plan(inverse_fft); delete[] spectrum; delete[] waveform; }
} InverseRealFFT;

typedef struct InverseComplexFFT {
  int fft_size;
  fft_complex *input;
  fft_complex *output;
  fft_plan inverse_fft;
  void initialize(int n) {
    fft_size = n;
    input = new fft_complex[n];
    output = new fft_complex[n];
    inverse_fft = fft_plan_dft_1d(n, input, output, FFT_BACKWARD, FFT_ESTIMATE);
  }
  void destroy() { fft_destroy_plan(inverse_fft); delete[] input; delete[] output; }
} InverseComplexFFT;

// minimum-phase spectrum from a logarithmic power spectrum (world_common.cpp:176-233): log_spectrum[0 .. n/2] in,
// minimum_phase_spectrum[0 .. n/2] out
typedef struct MinimumPhaseAnalysis {
  int fft_size;
  double *log_spectrum;
  fft_complex *minimum_phase_spectrum;
  fft_complex *cepstrum;
  fft_plan inverse_fft;
  fft_plan forward_fft;
  void initialize(int n) {
    fft_size = n;
    log_spectrum = new double[n];
    minimum_phase_spectrum = new fft_complex[n];
    cepstrum = new fft_complex[n];
    inverse_fft = fft_plan_dft_r2c_1d(n, log_spectrum, cepstrum, FFT_ESTIMATE);
    forward_fft = fft_plan_dft_1d(n, cepstrum, minimum_phase_spectrum, FFT_FORWARD, FFT_ESTIMATE);
  }
  void destroy() {
    fft_destroy_plan(forward_fft);
    fft_destroy_plan(inverse_fft);
    delete[] cepstrum;
    delete[] log_spectrum;
    delete[] minimum_phase_spectrum;
  }
  void compute() {
    const int half = fft_size / 2;
    for (int i = half + 1; i < fft_size; ++i) log_spectrum[i] = log_spectrum[fft_size - i];   // even extension
    fft_execute(inverse_fft);   // forward transform of a real, even sequence; conjugated below = the inverse one
    // fold the cepstrum onto non-negative quefrencies: ends once, interior twice, the rest zero
    cepstrum[0][1] = -cepstrum[0][1];
    for (int i = 1; i < half; ++i) { cepstrum[i][0] *= 2.0; cepstrum[i][1] *= -2.0; }
    cepstrum[half][1] = -cepstrum[half][1];
    for (int i = half + 1; i < fft_size; ++i) { cepstrum[i][0] = 0.0; cepstrum[i][1] = 0.0; }
    fft_execute(forward_fft);
    for (int i = 0; i <= half; ++i) {
      const double mag = std::exp(minimum_phase_spectrum[i][0] / fft_size);
      const double arg = minimum_phase_spectrum[i][1] / fft_size;
      minimum_phase_spectrum[i][0] = mag * std::cos(arg);
      minimum_phase_spectrum[i][1] = mag * std::sin(arg);
    }
  }
} MinimumPhaseAnalysis;

#endif
