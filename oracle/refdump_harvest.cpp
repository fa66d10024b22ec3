// TEST INFRASTRUCTURE: dumps Harvest's intermediate results from the UNMODIFIED reference
// sources.  The reference keeps every stage private, so this driver includes the class
// header with `private` re-defined and replays Harvest::generalBody
// (/root/reference/src/harvest.cpp:1380-1453) step by step, writing each intermediate as
// raw f64.  Used only while developing / debugging the CUDA stages and by
// tests/test_harvest_gpu.py to localise a mismatch; never linked into the product.
//
// usage: refdump_harvest x.f64 fs out_prefix [f0_floor=40] [f0_ceil=800]
#define private public
#include "harvest.hpp"
#undef private

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "world_constantnumbers.hpp"
#include "world_matlabfunctions.hpp"

using namespace world_class;

static void write_f64(const std::string &path, const double *p, size_t n) {
  FILE *fp = fopen(path.c_str(), "wb");
  fwrite(p, 8, n, fp);
  fclose(fp);
}
static void write_rows(const std::string &path, double **rows, int n_rows, int n_cols) {
  FILE *fp = fopen(path.c_str(), "wb");
  for (int i = 0; i < n_rows; ++i) fwrite(rows[i], 8, n_cols, fp);
  fclose(fp);
}

int main(int argc, char **argv) {
  if (argc < 4) return 2;
  FILE *fp = fopen(argv[1], "rb");
  fseek(fp, 0, SEEK_END);
  long n = ftell(fp) / 8;
  fseek(fp, 0, SEEK_SET);
  std::vector<double> x(n);
  if (fread(x.data(), 8, n, fp) != (size_t)n) return 2;
  fclose(fp);
  const int fs = atoi(argv[2]);
  const std::string out = argv[3];
  HarvestOption opt;
  opt.frame_period = 1.0;
  opt.f0_floor = argc > 4 ? atof(argv[4]) : 40.0;
  opt.f0_ceil = argc > 5 ? atof(argv[5]) : 800.0;
  Harvest h(fs, opt);

  // ---- replay of generalBody with frame_period = 1
  const int frame_period = 1;
  const double channels_in_octave = h.option_.channels_in_octave;
  h.x_ = x.data();
  h.x_length_ = (int)n;
  const int Lb = h.getSamples(fs, (int)n, frame_period);
  std::vector<double> tpos(Lb), f0(Lb);
  h.temporal_positions_ = tpos.data();
  double adjusted_f0_floor = h.option_.f0_floor * 0.9;
  double adjusted_f0_ceil = h.option_.f0_ceil * 1.1;
  int number_of_channels = 1 + static_cast<int>(log(adjusted_f0_ceil / adjusted_f0_floor) / world::kLog2 * channels_in_octave);
  std::vector<double> boundary_f0_list(number_of_channels);
  for (int i = 0; i < number_of_channels; ++i)
    boundary_f0_list[i] = adjusted_f0_floor * pow(2.0, static_cast<double>(i + 1) / channels_in_octave);
  h.y_length_ = (1 + static_cast<int>(h.x_length_ / h.decimation_ratio_));
  int fft_size = GetSuitableFFTSize(h.y_length_ + (4 * static_cast<int>(1.0 + h.actual_fs_ / boundary_f0_list[0] / 2.0)));
  h.y_ = new double[fft_size]();
  fft_complex *y_spectrum = new fft_complex[fft_size / 2 + 1];
  h.getWaveformAndSpectrum(fft_size, h.decimation_ratio_, y_spectrum);
  write_f64(out + ".y", h.y_, h.y_length_);
  h.f0_length_ = Lb;
  for (int i = 0; i < Lb; ++i) { tpos[i] = i * frame_period / 1000.0; f0[i] = 0.0; }
  int overlap_parameter = 7;
  int max_candidates = matlab_round(number_of_channels / 10) * overlap_parameter;
  h.f0_candidates_ = new double *[Lb];
  h.f0_candidates_score_ = new double *[Lb];
  for (int i = 0; i < Lb; ++i) {
    h.f0_candidates_[i] = new double[max_candidates]();
    h.f0_candidates_score_[i] = new double[max_candidates]();
  }
  // generalBodySub, opened up to dump the raw candidates
  double **raw = new double *[number_of_channels];
  for (int i = 0; i < number_of_channels; ++i) raw[i] = new double[Lb];
  h.getRawF0Candidates(boundary_f0_list.data(), number_of_channels, y_spectrum, fft_size, raw);
  write_rows(out + ".raw", raw, number_of_channels, Lb);
  int nc = h.detectOfficialF0Candidates(raw, number_of_channels, Lb, max_candidates, h.f0_candidates_);
  h.overlapF0Candidates(Lb, nc, h.f0_candidates_);
  h.number_of_candidates_ = nc * overlap_parameter;
  write_rows(out + ".cand0", h.f0_candidates_, Lb, max_candidates);
  h.refineF0Candidates();
  write_rows(out + ".cand1", h.f0_candidates_, Lb, max_candidates);
  write_rows(out + ".score1", h.f0_candidates_score_, Lb, max_candidates);
  h.removeUnreliableCandidates();
  write_rows(out + ".cand2", h.f0_candidates_, Lb, max_candidates);
  write_rows(out + ".score2", h.f0_candidates_score_, Lb, max_candidates);
  std::vector<double> c1(Lb), c2(Lb), best(Lb);
  h.searchF0Base(h.f0_candidates_, h.f0_candidates_score_, Lb, h.number_of_candidates_, c1.data());
  write_f64(out + ".base", c1.data(), Lb);
  h.fixStep1(c1.data(), 0.008, c2.data());
  write_f64(out + ".step1", c2.data(), Lb);
  h.fixStep2(c2.data(), 6, c1.data());
  write_f64(out + ".step2", c1.data(), Lb);
  h.fixStep3(c1.data(), 0.18, c2.data());
  write_f64(out + ".step3", c2.data(), Lb);
  h.fixStep4(c2.data(), 9, best.data());
  write_f64(out + ".step4", best.data(), Lb);
  h.smoothF0Contour(best.data(), f0.data());
  write_f64(out + ".f0", f0.data(), Lb);
  printf("{\"Lb\": %d, \"nch\": %d, \"max_candidates\": %d, \"nc\": %d, \"y_length\": %d, \"fft_size\": %d, \"r\": %d}\n",
         Lb, number_of_channels, max_candidates, nc, h.y_length_, fft_size, h.decimation_ratio_);
  return 0;  // leak everything: the destructor frees members we replaced
}
