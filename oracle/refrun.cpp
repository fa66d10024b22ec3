// TEST INFRASTRUCTURE: command-line driver over the UNMODIFIED reference classes
// (compiled from /root/reference/src/*.cpp by oracle/Makefile into oracle/_ref/).
// It mirrors the call sequence of /root/reference/test/test.cpp:76-264 (Harvest ->
// CheapTrick -> D4C -> Synthesis, then the codec round trip) but reads/writes raw
// little-endian f64 files so tests and bench.py can compare bytes.
//
// One process == one fresh randn() stream (the reference's RNG state is a function
// static, /root/reference/src/world_matlabfunctions.cpp:243-264), so every parity
// case runs this binary once.
//
// usage: refrun --in x.f64 --fs 48000 --out prefix [--stages hcdsk] [--f0-in f0.f64]
//               [--frame-period 5] [--harvest-f0-floor 40] [--harvest-f0-ceil 800]
//               [--ct-f0-floor 71] [--d4c-threshold 0.85] [--codec-nd 60]
//               [--repeat 1] [--no-write]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "harvest.hpp"
#include "cheaptrick.hpp"
#include "d4c.hpp"
#include "synthesis.hpp"
#include "codec.hpp"

using namespace world_class;

static std::vector<double> read_f64(const std::string &path) {
  FILE *fp = fopen(path.c_str(), "rb");
  if (!fp) { fprintf(stderr, "refrun: cannot open %s\n", path.c_str()); exit(2); }
  fseek(fp, 0, SEEK_END);
  long n = ftell(fp) / 8;
  fseek(fp, 0, SEEK_SET);
  std::vector<double> v(n);
  if (fread(v.data(), 8, n, fp) != (size_t)n) { fprintf(stderr, "refrun: short read\n"); exit(2); }
  fclose(fp);
  return v;
}

static void write_f64(const std::string &path, const double *p, size_t n) {
  FILE *fp = fopen(path.c_str(), "wb");
  if (!fp) { fprintf(stderr, "refrun: cannot write %s\n", path.c_str()); exit(2); }
  fwrite(p, 8, n, fp);
  fclose(fp);
}

static void write_rows(const std::string &path, double **rows, int n_rows, int n_cols) {
  FILE *fp = fopen(path.c_str(), "wb");
  if (!fp) { fprintf(stderr, "refrun: cannot write %s\n", path.c_str()); exit(2); }
  for (int i = 0; i < n_rows; ++i) fwrite(rows[i], 8, n_cols, fp);
  fclose(fp);
}

static double **alloc_rows(int n_rows, int n_cols) {
  double **r = new double *[n_rows];
  for (int i = 0; i < n_rows; ++i) r[i] = new double[n_cols];
  return r;
}

static void free_rows(double **r, int n_rows) {
  for (int i = 0; i < n_rows; ++i) delete[] r[i];
  delete[] r;
}

typedef std::chrono::steady_clock clk;
static double ms_since(clk::time_point t0) {
  return std::chrono::duration<double, std::milli>(clk::now() - t0).count();
}

int main(int argc, char **argv) {
  std::string in_path, out_prefix, f0_in, stages = "hcds";
  int fs = 0, codec_nd = 60, repeat = 1;
  bool no_write = false;
  double frame_period = 5.0, h_floor = 40.0, h_ceil = 800.0, ct_floor = 71.0, d4c_thr = 0.85;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto next = [&]() -> const char * {
      if (i + 1 >= argc) { fprintf(stderr, "refrun: missing value for %s\n", a.c_str()); exit(2); }
      return argv[++i];
    };
    if (a == "--in") in_path = next();
    else if (a == "--out") out_prefix = next();
    else if (a == "--f0-in") f0_in = next();
    else if (a == "--stages") stages = next();
    else if (a == "--fs") fs = atoi(next());
    else if (a == "--frame-period") frame_period = atof(next());
    else if (a == "--harvest-f0-floor") h_floor = atof(next());
    else if (a == "--harvest-f0-ceil") h_ceil = atof(next());
    else if (a == "--ct-f0-floor") ct_floor = atof(next());
    else if (a == "--d4c-threshold") d4c_thr = atof(next());
    else if (a == "--codec-nd") codec_nd = atoi(next());
    else if (a == "--repeat") repeat = atoi(next());
    else if (a == "--no-write") no_write = true;
    else { fprintf(stderr, "refrun: unknown arg %s\n", a.c_str()); return 2; }
  }
  if (in_path.empty() || fs <= 0 || (out_prefix.empty() && !no_write)) {
    fprintf(stderr, "refrun: need --in, --fs, --out\n");
    return 2;
  }
  const bool do_h = stages.find('h') != std::string::npos;
  const bool do_c = stages.find('c') != std::string::npos;
  const bool do_d = stages.find('d') != std::string::npos;
  const bool do_s = stages.find('s') != std::string::npos;
  const bool do_k = stages.find('k') != std::string::npos;

  std::vector<double> x = read_f64(in_path);
  const int x_length = (int)x.size();
  int threads = 1;
#ifdef _OPENMP
  threads = omp_get_num_procs();
#endif

  for (int rep = 0; rep < repeat; ++rep) {
    double t_h = 0, t_c = 0, t_d = 0, t_s = 0, t_k = 0;

    // ---- Harvest (test.cpp:76-114)
    HarvestOption hopt;
    hopt.frame_period = frame_period;
    hopt.f0_floor = h_floor;
    hopt.f0_ceil = h_ceil;
    Harvest *harvest = new Harvest(fs, hopt);
    const int f0_length = harvest->getSamples(fs, x_length);
    std::vector<double> f0(f0_length), tpos(f0_length);
    if (do_h) {
      auto t0 = clk::now();
      harvest->compute(x.data(), x_length, tpos.data(), f0.data());
      t_h = ms_since(t0);
    } else {
      std::vector<double> f0_file = read_f64(f0_in);
      if ((int)f0_file.size() != f0_length) {
        fprintf(stderr, "refrun: f0 file has %zu entries, expected %d\n", f0_file.size(), f0_length);
        return 2;
      }
      f0 = f0_file;
      for (int i = 0; i < f0_length; ++i) tpos[i] = i * frame_period / 1000.0;
    }
    delete harvest;

    // ---- CheapTrick (test.cpp:117-161)
    CheapTrickOption copt;
    copt.f0_floor = ct_floor;
    CheapTrick *cheaptrick = new CheapTrick(fs, copt);
    const int fft_size = cheaptrick->getFFTSizeForCheapTrick(fs, copt.f0_floor);
    const int bins = fft_size / 2 + 1;
    double **sp = alloc_rows(f0_length, bins);
    double **ap = alloc_rows(f0_length, bins);
    if (do_c) {
      auto t0 = clk::now();
      cheaptrick->compute(x.data(), x_length, tpos.data(), f0.data(), f0_length, sp);
      t_c = ms_since(t0);
    }
    delete cheaptrick;

    // ---- D4C (test.cpp:164-198)
    if (do_d) {
      D4COption dopt;
      dopt.threshold = d4c_thr;
      D4C *d4c = new D4C(fs, dopt);
      auto t0 = clk::now();
      d4c->compute(x.data(), x_length, tpos.data(), f0.data(), f0_length, fft_size, ap);
      t_d = ms_since(t0);
      delete d4c;
    }

    // ---- Synthesis (test.cpp:245-264, y_length from :362-363)
    const int y_length = static_cast<int>((f0_length - 1) * frame_period / 1000.0 * fs) + 1;
    std::vector<double> y(y_length, 0.0);
    if (do_s) {
      Synthesis *synthesis = new Synthesis(fs, fft_size, frame_period);
      auto t0 = clk::now();
      synthesis->compute(f0.data(), f0_length, sp, ap, y_length, y.data());
      t_s = ms_since(t0);
      delete synthesis;
    }

    // ---- codec round trip (codec.hpp:23-88); not exercised by test.cpp
    const int n_ap = GetNumberOfAperiodicities(fs);
    double **csp = NULL, **cap = NULL, **dsp = NULL, **dap = NULL;
    if (do_k) {
      csp = alloc_rows(f0_length, codec_nd);
      cap = alloc_rows(f0_length, n_ap > 0 ? n_ap : 1);
      dsp = alloc_rows(f0_length, bins);
      dap = alloc_rows(f0_length, bins);
      auto t0 = clk::now();
      CodeSpectralEnvelope(sp, f0_length, fs, fft_size, codec_nd, csp);
      CodeAperiodicity(ap, f0_length, fs, fft_size, cap);
      DecodeSpectralEnvelope(csp, f0_length, fs, fft_size, codec_nd, dsp);
      DecodeAperiodicity(cap, f0_length, fs, fft_size, dap);
      t_k = ms_since(t0);
    }

    printf("{\"rep\": %d, \"threads\": %d, \"fs\": %d, \"x_length\": %d, \"f0_length\": %d, "
           "\"fft_size\": %d, \"y_length\": %d, \"n_ap\": %d, \"harvest_ms\": %.4f, "
           "\"cheaptrick_ms\": %.4f, \"d4c_ms\": %.4f, \"synthesis_ms\": %.4f, \"codec_ms\": %.4f}\n",
           rep, threads, fs, x_length, f0_length, fft_size, y_length, n_ap, t_h, t_c, t_d, t_s, t_k);
    fflush(stdout);

    if (rep == 0 && !no_write) {
      write_f64(out_prefix + ".tpos", tpos.data(), f0_length);
      write_f64(out_prefix + ".f0", f0.data(), f0_length);
      if (do_c) write_rows(out_prefix + ".sp", sp, f0_length, bins);
      if (do_d) write_rows(out_prefix + ".ap", ap, f0_length, bins);
      if (do_s) write_f64(out_prefix + ".y", y.data(), y_length);
      if (do_k) {
        write_rows(out_prefix + ".csp", csp, f0_length, codec_nd);
        write_rows(out_prefix + ".cap", cap, f0_length, n_ap);
        write_rows(out_prefix + ".dsp", dsp, f0_length, bins);
        write_rows(out_prefix + ".dap", dap, f0_length, bins);
      }
    }
    free_rows(sp, f0_length);
    free_rows(ap, f0_length);
    if (do_k) {
      free_rows(csp, f0_length); free_rows(cap, f0_length);
      free_rows(dsp, f0_length); free_rows(dap, f0_length);
    }
  }
  return 0;
}
