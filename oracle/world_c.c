/* TEST INFRASTRUCTURE -- plain C restatement of the two strictly sequential pieces of the reference
 * that a numpy oracle cannot vectorise: the randn() generator and the running phase sum.
 * Compiled by __graft_entry__.build() / oracle/world_np.py into oracle/libworldoracle.so.
 * Never linked into or called by the product. */
#include <math.h>
#include <stdint.h>

/* /root/reference/src/world_matlabfunctions.cpp:243-264 */
void oracle_randn_fill(uint32_t state[4], long n, double *out) {
  uint32_t x = state[0], y = state[1], z = state[2], w = state[3];
  for (long k = 0; k < n; ++k) {
    uint32_t t;
    x = y; y = z; z = w;              /* partial shift; the first t of the reference is dead */
    uint32_t tmp = 0;
    for (int i = 0; i < 12; ++i) {
      t = x ^ (x << 11);
      x = y; y = z; z = w;
      w = (w ^ (w >> 19)) ^ (t ^ (t >> 8));
      tmp += w >> 4;
    }
    out[k] = tmp / 268435456.0 - 6.0;
  }
  state[0] = x; state[1] = y; state[2] = z; state[3] = w;
}

/* /root/reference/src/synthesis.cpp:245-288 (getPulseLocationsForTimeBase); returns number of pulses */
int oracle_pulse_locations(const double *interpolated_f0, int y_length, int fs, int *pulse_index,
                           double *pulse_time_shift) {
  const double kPi = 3.1415926535897932384;
  const double two_pi = 2.0 * kPi;
  const double const_val = two_pi / fs;
  double total = interpolated_f0[0] * const_val;
  double wrap_prev = fmod(total, two_pi);
  int n = 0;
  for (int ii = 1; ii < y_length; ++ii) {
    total = total + interpolated_f0[ii] * const_val;
    const double wrap = fmod(total, two_pi);
    if (fabs(wrap - wrap_prev) > kPi) {
      const double y1 = wrap_prev - two_pi;
      const double y2 = wrap;
      const double xx = -y1 / (y2 - y1);
      pulse_index[n] = ii - 1;
      pulse_time_shift[n] = xx / fs;
      ++n;
    }
    wrap_prev = wrap;
  }
  return n;
}
