// TEST INFRASTRUCTURE: runs the reference demo's ParameterModification (test/test.cpp:201-243) on raw f64
// arrays.  The function lives in an anonymous namespace of test.cpp, so that file is included into this
// translation unit where it lies (its main() renamed); nothing of it is copied.
//
//   refmod <f0.f64> <sp.f64> <fs> <fft_size> <f0_length> <shift|-> <ratio|-> <out prefix>
//
// "-" leaves the corresponding argument out, exactly like a shorter command line of the demo does
// (argc >= 4: F0 scaling, argc >= 5: spectral stretching).
#define main reference_test_main
#include REF_TEST_CPP
#undef main

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static bool read_f64(const char *path, std::vector<double> &v, size_t n) {
  FILE *f = fopen(path, "rb");
  if (!f) return false;
  v.resize(n);
  const size_t got = fread(v.data(), sizeof(double), n, f);
  fclose(f);
  return got == n;
}

int main(int argc, char *argv[]) {
  if (argc != 9) { fprintf(stderr, "usage: refmod f0 sp fs fft_size f0_length shift ratio out\n"); return 2; }
  const int fs = atoi(argv[3]), fft_size = atoi(argv[4]), L = atoi(argv[5]);
  const int bins = fft_size / 2 + 1;
  std::vector<double> f0, sp;
  if (!read_f64(argv[1], f0, L) || !read_f64(argv[2], sp, (size_t)L * bins)) { fprintf(stderr, "short input\n"); return 1; }
  std::vector<double *> rows(L);
  // rows as long as the demo allocates them (test/test.cpp:146-149): fft_size / 2 + 1 entries
  for (int i = 0; i < L; ++i) rows[i] = sp.data() + (size_t)i * bins;
  // the demo's command line: argv[3] = shift, argv[4] = ratio
  char a0[] = "test", a1[] = "in.wav", a2[] = "out";
  std::vector<char *> av = {a0, a1, a2};
  if (strcmp(argv[6], "-") != 0) av.push_back(argv[6]);
  if (strcmp(argv[7], "-") != 0) {
    if (av.size() < 4) { fprintf(stderr, "ratio needs a shift\n"); return 2; }
    av.push_back(argv[7]);
  }
  ParameterModification((int)av.size(), av.data(), fs, L, fft_size, f0.data(), rows.data());
  std::string out(argv[8]);
  FILE *f = fopen((out + ".f0").c_str(), "wb");
  fwrite(f0.data(), sizeof(double), L, f);
  fclose(f);
  f = fopen((out + ".sp").c_str(), "wb");
  fwrite(sp.data(), sizeof(double), (size_t)L * bins, f);
  fclose(f);
  return 0;
}
