// TEST INFRASTRUCTURE: drives the reference's own file I/O (tools/parameterio.cpp, tools/audioio.cpp, compiled
// where they lie by oracle/Makefile) on raw f64 arrays, so that the library's writers can be compared byte for
// byte with the reference's and its readers value for value.
//
//   refio write <dir> <fs> <f0_length> <fft_size> <number_of_dimensions> <frame_period> <x_length>
//       reads  <dir>/{tpos,f0,sp,ap,x}.f64 (sp/ap: f0_length rows of nd' doubles, nd' = nd ? nd : fft_size/2+1)
//       writes <dir>/ref_f0.bin ref_f0.txt ref_sp.bin ref_ap.bin ref_x.wav
//   refio read <dir> <prefix>
//       reads  <dir>/<prefix>_f0.bin _sp.bin _ap.bin _x.wav with the reference's readers
//       writes <dir>/<prefix>_{tpos,f0,sp,ap,x}.rd.f64 and prints one JSON line with the header fields
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "audioio.hpp"
#include "parameterio.hpp"

static std::vector<double> slurp(const std::string &path, size_t n) {
  std::vector<double> v(n);
  FILE *f = fopen(path.c_str(), "rb");
  if (!f || fread(v.data(), 8, n, f) != n) { fprintf(stderr, "short read %s\n", path.c_str()); exit(1); }
  fclose(f);
  return v;
}
static void dump(const std::string &path, const double *v, size_t n) {
  FILE *f = fopen(path.c_str(), "wb");
  fwrite(v, 8, n, f);
  fclose(f);
}

int main(int argc, char **argv) {
  if (argc >= 9 && !strcmp(argv[1], "write")) {
    const std::string d = argv[2];
    const int fs = atoi(argv[3]), L = atoi(argv[4]), fft = atoi(argv[5]), nd = atoi(argv[6]);
    const double fp = atof(argv[7]);
    const int nx = atoi(argv[8]);
    const int dims = nd ? nd : fft / 2 + 1;
    std::vector<double> tpos = slurp(d + "/tpos.f64", L), f0 = slurp(d + "/f0.f64", L);
    std::vector<double> sp = slurp(d + "/sp.f64", (size_t)L * dims), ap = slurp(d + "/ap.f64", (size_t)L * dims);
    std::vector<double> x = slurp(d + "/x.f64", nx);
    std::vector<double *> rs(L), ra(L);
    for (int i = 0; i < L; ++i) { rs[i] = sp.data() + (size_t)i * dims; ra[i] = ap.data() + (size_t)i * dims; }
    WriteF0((d + "/ref_f0.bin").c_str(), L, fp, tpos.data(), f0.data(), 0);
    WriteF0((d + "/ref_f0.txt").c_str(), L, fp, tpos.data(), f0.data(), 1);
    WriteSpectralEnvelope((d + "/ref_sp.bin").c_str(), fs, L, fp, fft, nd, rs.data());
    WriteAperiodicity((d + "/ref_ap.bin").c_str(), fs, L, fp, fft, nd, ra.data());
    wavwrite(x.data(), nx, fs, 16, (d + "/ref_x.wav").c_str());
    return 0;
  }
  if (argc >= 4 && !strcmp(argv[1], "read")) {
    const std::string d = argv[2], p = d + "/" + argv[3];
    const std::string f0f = p + "_f0.bin", spf = p + "_sp.bin", apf = p + "_ap.bin", wav = p + "_x.wav";
    const int L = (int)GetHeaderInformation(f0f.c_str(), "NOF ");
    const double fp = GetHeaderInformation(spf.c_str(), "FP  ");
    const int fft = (int)GetHeaderInformation(spf.c_str(), "FFT ");
    const int nd = (int)GetHeaderInformation(spf.c_str(), "NOD ");
    const int fsh = (int)GetHeaderInformation(apf.c_str(), "FS  ");
    const int dims = nd ? nd : fft / 2 + 1;
    std::vector<double> tpos(L), f0(L), sp((size_t)L * dims), ap((size_t)L * dims);
    std::vector<double *> rs(L), ra(L);
    for (int i = 0; i < L; ++i) { rs[i] = sp.data() + (size_t)i * dims; ra[i] = ap.data() + (size_t)i * dims; }
    const int ok_f0 = ReadF0(f0f.c_str(), tpos.data(), f0.data());
    const int ok_sp = ReadSpectralEnvelope(spf.c_str(), rs.data());
    const int ok_ap = ReadAperiodicity(apf.c_str(), ra.data());
    const int nx = GetAudioLength(wav.c_str());
    int fs = 0, nbit = 0;
    std::vector<double> x(nx > 0 ? nx : 0);
    if (nx > 0) wavread(wav.c_str(), &fs, &nbit, x.data());
    dump(p + "_tpos.rd.f64", tpos.data(), L);
    dump(p + "_f0.rd.f64", f0.data(), L);
    dump(p + "_sp.rd.f64", sp.data(), sp.size());
    dump(p + "_ap.rd.f64", ap.data(), ap.size());
    dump(p + "_x.rd.f64", x.data(), x.size());
    printf("{\"NOF\": %d, \"FP\": %.17g, \"FFT\": %d, \"NOD\": %d, \"FS\": %d, \"ok\": [%d, %d, %d], \"wav_length\": %d, \"wav_fs\": %d, \"wav_nbit\": %d}\n",
           L, fp, fft, nd, fsh, ok_f0, ok_sp, ok_ap, nx, fs, nbit);
    return 0;
  }
  fprintf(stderr, "usage: refio write|read ...\n");
  return 2;
}
