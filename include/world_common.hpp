// Source-compatibility header (worldb200) for the parts of /root/reference/include/world_common.hpp
// that callers of the class API touch: the min/max helpers, GetSuitableFFTSize and
// GetSafeAperiodicity.  The FFT plan structs of the reference are internal scratch of its CPU
// stages and have no counterpart here (every transform lives in GPU shared memory).
#ifndef WORLD_COMMON_HPP
#define WORLD_COMMON_HPP

#include <cmath>

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

#endif
