// The public C-linkage helpers of world_common.hpp / world_matlabfunctions.hpp / world_fft.hpp, driven on seeded
// inputs.  ONE source, built twice: against the reference's headers + sources (oracle/_ref/refhelpers, by
// oracle/Makefile) and against this repository's drop-in headers + libworldb200.so (tests/dropin/_bin/helpers_main, by
// __graft_entry__.build()); the tests compare the two outputs.
//
//   helpers_main host|all <out.f64>      host: header-only helpers (no GPU needed); all: + decimate, the FFT structs
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "world_common.hpp"
#include "world_matlabfunctions.hpp"

static unsigned long long g_state = 88172645463325252ull;
static double uniform01() {   // xorshift64: the same inputs in both builds
  g_state ^= g_state << 13; g_state ^= g_state >> 7; g_state ^= g_state << 17;
  return (double)(g_state >> 11) / 9007199254740992.0;
}

static std::vector<double> g_out;
static void put(const double *v, int n) { g_out.insert(g_out.end(), v, v + n); }
static void put(const std::vector<double> &v) { g_out.insert(g_out.end(), v.begin(), v.end()); }

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  const bool all = strcmp(argv[1], "all") == 0;
  const int fs = 48000, fft_size = 2048, half = fft_size / 2;

  // histc / interp1 (with extrapolation beyond both ends) / interp1Q
  {
    const int n = 40, m = 97;
    std::vector<double> x(n), y(n), xi(m), yi(m), hidx(m);
    double t = -3.0;
    for (int i = 0; i < n; ++i) { t += 0.05 + uniform01(); x[i] = t; y[i] = uniform01() * 4.0 - 2.0; }
    for (int i = 0; i < m; ++i) xi[i] = x[0] - 1.5 + (x[n - 1] - x[0] + 3.0) * i / (m - 1);
    xi[10] = x[3]; xi[11] = x[n - 1]; xi[12] = x[0];   // queries on knots
    std::vector<int> index(m);
    histc(x.data(), n, xi.data(), m, index.data());
    for (int i = 0; i < m; ++i) hidx[i] = index[i];
    put(hidx);
    interp1(x.data(), y.data(), n, xi.data(), m, yi.data());
    put(yi);
    std::vector<double> q(m), qi(m);
    for (int i = 0; i < m; ++i) q[i] = 100.0 + 7.5 * (n - 1.000001) * i / (m - 1);
    interp1Q(100.0, 7.5, y.data(), n, q.data(), m, qi.data());
    put(qi);
    std::vector<double> sh(n), df(n - 1);
    fftshift(y.data(), n, sh.data());
    diff(y.data(), n, df.data());
    put(sh); put(df);
    double r[6] = {(double)matlab_round(2.5), (double)matlab_round(-2.5), (double)matlab_round(0.49999), (double)matlab_round(-0.5),
                   (double)GetSuitableFFTSize(1025), GetSafeAperiodicity(1.5) + GetSafeAperiodicity(-1.0)};
    put(r, 6);
  }
  // DCCorrection / LinearSmoothing / NuttallWindow
  {
    std::vector<double> spec(half + 1), out(half + 1);
    for (int i = 0; i <= half; ++i) spec[i] = std::exp(uniform01() * 8.0 - 6.0);
    const double f0s[3] = {71.3, 140.0, 612.5};
    for (int k = 0; k < 3; ++k) {
      out = spec;
      DCCorrection(spec.data(), f0s[k], fs, fft_size, out.data());
      put(out);
      LinearSmoothing(spec.data(), f0s[k] * 2.0 / 3.0, fs, fft_size, out.data());
      put(out);
    }
    std::vector<double> w(513);
    NuttallWindow(513, w.data());
    put(w);
  }
  if (all) {
    // decimate
    const int rs[3] = {2, 6, 11};
    for (int k = 0; k < 3; ++k) {
      const int n = 4000 + 37 * k, r = rs[k];
      std::vector<double> x(n), y(n / r + 8, 0.0);
      for (int i = 0; i < n; ++i) x[i] = std::sin(0.01 * i * (k + 1)) * 0.5 + (uniform01() - 0.5) * 0.2;
      decimate(x.data(), n, r, y.data());
      const int nout = n / r + 1, nbeg = r - r * nout + n;
      put(y.data(), (n + 9 - nbeg + r - 1) / r);
    }
    // the FFT structs
    {
      ForwardRealFFT f;
      f.initialize(fft_size);
      for (int i = 0; i < fft_size; ++i) f.waveform[i] = uniform01() - 0.5;
      fft_execute(f.forward_fft);
      put(&f.spectrum[0][0], 2 * (half + 1));
      InverseRealFFT inv;
      inv.initialize(fft_size);
      for (int i = 0; i <= half; ++i) { inv.spectrum[i][0] = f.spectrum[i][0]; inv.spectrum[i][1] = f.spectrum[i][1]; }
      fft_execute(inv.inverse_fft);
      put(inv.waveform, fft_size);
      InverseComplexFFT c;
      c.initialize(fft_size);
      for (int i = 0; i < fft_size; ++i) { c.input[i][0] = uniform01() - 0.5; c.input[i][1] = uniform01() - 0.5; }
      fft_execute(c.inverse_fft);
      put(&c.output[0][0], 2 * fft_size);
      MinimumPhaseAnalysis mp;
      mp.initialize(fft_size);
      for (int i = 0; i <= half; ++i) mp.log_spectrum[i] = std::log(std::exp(uniform01() * 6.0 - 5.0)) / 2.0;
      mp.compute();
      put(&mp.minimum_phase_spectrum[0][0], 2 * (half + 1));
      f.destroy(); inv.destroy(); c.destroy(); mp.destroy();
    }
  }
  FILE *fp = fopen(argv[2], "wb");
  if (!fp) return 3;
  const bool ok = fwrite(g_out.data(), sizeof(double), g_out.size(), fp) == g_out.size();
  fclose(fp);
  printf("%zu values %s\n", g_out.size(), ok ? "ok" : "WRITE FAILED");
  return ok ? 0 : 4;
}
