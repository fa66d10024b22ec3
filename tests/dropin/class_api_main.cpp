// A consumer of the drop-in C++ headers, written the way the reference's own demo uses its classes
// (/root/reference/test/test.cpp:76-264): separately allocated rows, copy-initialised stage objects, options set
// field by field, interp1 between analysis and synthesis, the codec free functions on the results.  Raw float64
// in / out so that the GPU test can compare every array with the reference process (oracle/_ref/refrun).
//
//   class_api_main <x.f64> <fs> <out prefix> [f0 shift] [formant ratio]
//
// Built by __graft_entry__.build() against include/*.hpp and libworldb200.so only.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cheaptrick.hpp"
#include "codec.hpp"
#include "d4c.hpp"
#include "harvest.hpp"
#include "synthesis.hpp"
#include "world_common.hpp"
#include "world_matlabfunctions.hpp"

using namespace world_class;

static bool write_f64(const char *prefix, const char *name, const double *v, size_t n) {
  char path[1024];
  snprintf(path, sizeof(path), "%s_%s.f64", prefix, name);
  FILE *f = fopen(path, "wb");
  if (!f) return false;
  const bool ok = fwrite(v, sizeof(double), n, f) == n;
  fclose(f);
  return ok;
}

static bool write_rows(const char *prefix, const char *name, double *const *rows, int n_rows, int cols) {
  std::vector<double> flat((size_t)n_rows * cols);
  for (int i = 0; i < n_rows; ++i)
    for (int j = 0; j < cols; ++j) flat[(size_t)i * cols + j] = rows[i][j];
  return write_f64(prefix, name, flat.data(), flat.size());
}

int main(int argc, char **argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s x.f64 fs out_prefix [f0_shift] [ratio]\n", argv[0]);
    return 2;
  }
  const int fs = atoi(argv[2]);
  const char *prefix = argv[3];
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 3;
  fseek(f, 0, SEEK_END);
  const int x_length = (int)(ftell(f) / sizeof(double));
  fseek(f, 0, SEEK_SET);
  double *x = new double[x_length];
  if (fread(x, sizeof(double), x_length, f) != (size_t)x_length) return 3;
  fclose(f);
  const double frame_period = 5.0;

  // ---- Harvest (test.cpp:76-113)
  HarvestOption h_option;
  h_option.frame_period = frame_period;
  h_option.f0_floor = 40.0;
  Harvest harvest = Harvest(fs, h_option);
  const int f0_length = harvest.getSamples(fs, x_length);
  double *f0 = new double[f0_length];
  double *time_axis = new double[f0_length];
  harvest.compute(x, x_length, time_axis, f0);

  // ---- CheapTrick (test.cpp:116-160)
  CheapTrickOption c_option;
  c_option.f0_floor = 71.0;
  CheapTrick cheaptrick = CheapTrick(fs, c_option);
  const int fft_size = cheaptrick.getFFTSizeForCheapTrick(fs, c_option.f0_floor);
  const int bins = fft_size / 2 + 1;
  double **spectrogram = new double *[f0_length];
  for (int i = 0; i < f0_length; ++i) spectrogram[i] = new double[bins];
  cheaptrick.compute(x, x_length, time_axis, f0, f0_length, spectrogram);

  // ---- D4C (test.cpp:163-198)
  double **aperiodicity = new double *[f0_length];
  for (int i = 0; i < f0_length; ++i) aperiodicity[i] = new double[bins];
  D4COption d_option;
  d_option.threshold = 0.85;
  D4C d4c = D4C(fs, d_option);
  d4c.compute(x, x_length, time_axis, f0, f0_length, fft_size, aperiodicity);

  bool ok = write_f64(prefix, "tpos", time_axis, f0_length) && write_f64(prefix, "f0", f0, f0_length) &&
            write_rows(prefix, "sp", spectrogram, f0_length, bins) && write_rows(prefix, "ap", aperiodicity, f0_length, bins);

  // ---- codec round trip on the analysis results (codec.hpp:23-88)
  const int nd = 60;
  const int n_ap = GetNumberOfAperiodicities(fs);
  double **coded_sp = new double *[f0_length], **coded_ap = new double *[f0_length];
  for (int i = 0; i < f0_length; ++i) { coded_sp[i] = new double[nd]; coded_ap[i] = new double[n_ap > 0 ? n_ap : 1]; }
  CodeSpectralEnvelope(spectrogram, f0_length, fs, fft_size, nd, coded_sp);
  if (n_ap > 0) CodeAperiodicity(aperiodicity, f0_length, fs, fft_size, coded_ap);
  ok = ok && write_rows(prefix, "coded_sp", coded_sp, f0_length, nd);
  if (n_ap > 0) ok = ok && write_rows(prefix, "coded_ap", coded_ap, f0_length, n_ap);

  // ---- ParameterModification as the demo writes it (test.cpp:201-243): F0 scaling, spectral stretching by interp1
  if (argc > 4) {
    const double shift = atof(argv[4]);
    for (int i = 0; i < f0_length; ++i) f0[i] *= shift;
  }
  if (argc > 5) {
    const double ratio = atof(argv[5]);
    double *freq_axis1 = new double[fft_size], *freq_axis2 = new double[fft_size];
    double *spectrum1 = new double[fft_size], *spectrum2 = new double[fft_size];
    for (int i = 0; i <= fft_size / 2; ++i) {
      freq_axis1[i] = ratio * i / fft_size * fs;
      freq_axis2[i] = static_cast<double>(i) / fft_size * fs;
    }
    for (int i = 0; i < f0_length; ++i) {
      for (int j = 0; j <= fft_size / 2; ++j) spectrum1[j] = log(spectrogram[i][j]);
      interp1(freq_axis1, spectrum1, fft_size / 2 + 1, freq_axis2, fft_size / 2 + 1, spectrum2);
      for (int j = 0; j <= fft_size / 2; ++j) spectrogram[i][j] = exp(spectrum2[j]);
      if (ratio >= 1.0) continue;
      for (int j = static_cast<int>(fft_size / 2.0 * ratio); j <= fft_size / 2; ++j)
        spectrogram[i][j] = spectrogram[i][static_cast<int>(fft_size / 2.0 * ratio) - 1];
    }
    delete[] spectrum1; delete[] spectrum2; delete[] freq_axis1; delete[] freq_axis2;
  }

  // ---- Synthesis (test.cpp:246-264, :362-368)
  const int y_length = static_cast<int>((f0_length - 1) * frame_period / 1000.0 * fs) + 1;
  double *y = new double[y_length]();
  Synthesis synthesis = Synthesis(fs, fft_size, frame_period);
  synthesis.compute(f0, f0_length, spectrogram, aperiodicity, y_length, y);
  ok = ok && write_f64(prefix, "y", y, y_length) && write_f64(prefix, "f0_syn", f0, f0_length);

  printf("frames %d fft_size %d samples %d %s\n", f0_length, fft_size, y_length, ok ? "ok" : "WRITE FAILED");
  for (int i = 0; i < f0_length; ++i) { delete[] spectrogram[i]; delete[] aperiodicity[i]; delete[] coded_sp[i]; delete[] coded_ap[i]; }
  delete[] spectrogram; delete[] aperiodicity; delete[] coded_sp; delete[] coded_ap;
  delete[] f0; delete[] time_axis; delete[] x; delete[] y;
  return ok ? 0 : 4;
}
