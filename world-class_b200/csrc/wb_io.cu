// Data formats either side of the hot path:
//
//  * the parameter files of the reference's tooling (tools/parameterio.cpp:60-244: "F0  " / "SPEC" / "AP  "
//    containers with NOF / FP / FFT / NOD / FS header fields) and its minimal RIFF reader / writer
//    (tools/audioio.cpp:121-253), byte-compatible in both directions.  Host code (the reference's is
//    compiled host code too); the matrix variants take the contiguous [frames][dims] arrays the device
//    path produces as well as the reference's separately allocated rows;
//  * the sample-format conversions of that reader / writer as device kernels, so that a wav -> wav run moves
//    16-bit PCM over PCIe instead of fp64 (4x fewer bytes each way), and fp64 -> fp32 narrowing of the
//    analysis outputs for consumers that do not want doubles.
#include "../../include/worldb200.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "wb_internal.h"

namespace {

// ---- parameter containers (tools/parameterio.cpp) ------------------------------------------------
// layout: 4-char magic, then tagged fields: 4-char tag + int32 (NOF, FFT, NOD, FS) or + float64 (FP)

struct FileCloser {
  FILE *fp;
  explicit FileCloser(FILE *f) : fp(f) {}
  ~FileCloser() { if (fp) fclose(fp); }
};

bool put_tag_i32(FILE *fp, const char *tag, double value) {  // the reference converts through double (parameterio.cpp:17-21)
  const int32_t v = static_cast<int32_t>(value);
  return fwrite(tag, 1, 4, fp) == 4 && fwrite(&v, 4, 1, fp) == 1;
}
bool put_tag_f64(FILE *fp, const char *tag, double value) {
  return fwrite(tag, 1, 4, fp) == 4 && fwrite(&value, 8, 1, fp) == 1;
}
bool expect_magic(FILE *fp, const char *magic) {
  char got[4];
  return fread(got, 1, 4, fp) == 4 && memcmp(got, magic, 4) == 0;
}

// header of SPEC / AP after the magic (parameterio.cpp:28-46); NOD == 0 means fft_size / 2 + 1
int read_matrix_header(FILE *fp, int *frames, int *fft_size, int *dims) {
  unsigned char h[44];
  if (fread(h, 1, 44, fp) != 44) return WB_ERR_ARG;
  int32_t nof, fft, nod;
  memcpy(&nof, h + 4, 4);    // "NOF " i32 | "FP  " f64 | "FFT " i32 | "NOD " i32 | "FS  " i32
  memcpy(&fft, h + 24, 4);
  memcpy(&nod, h + 32, 4);
  *frames = nof;
  *fft_size = fft;
  *dims = nod == 0 ? fft / 2 + 1 : nod;
  return (nof < 0 || *dims <= 0) ? WB_ERR_ARG : WB_OK;
}

int write_matrix(const char *magic, const char *filename, int fs, int f0_length, double frame_period, int fft_size,
                 int number_of_dimensions, const double *const *rows, const double *contiguous, long long row_stride) {
  if (!filename || f0_length < 0 || (!rows && !contiguous && f0_length > 0)) return WB_ERR_ARG;
  FILE *fp = fopen(filename, "wb");
  if (!fp) return WB_ERR_ARG;
  FileCloser guard(fp);
  bool ok = fwrite(magic, 1, 4, fp) == 4 && put_tag_i32(fp, "NOF ", f0_length) && put_tag_f64(fp, "FP  ", frame_period) &&
            put_tag_i32(fp, "FFT ", fft_size) && put_tag_i32(fp, "NOD ", number_of_dimensions) && put_tag_i32(fp, "FS  ", fs);
  const int dims = number_of_dimensions == 0 ? fft_size / 2 + 1 : number_of_dimensions;  // parameterio.cpp:170-171
  if (dims <= 0) return WB_ERR_ARG;
  for (int i = 0; ok && i < f0_length; ++i) {
    const double *row = rows ? rows[i] : contiguous + (size_t)i * row_stride;
    ok = fwrite(row, 8, dims, fp) == (size_t)dims;
  }
  return ok ? WB_OK : WB_ERR_ARG;
}

int read_matrix(const char *magic, const char *filename, double **rows, double *contiguous, long long row_stride) {
  if (!filename || (!rows && !contiguous)) return WB_ERR_ARG;
  FILE *fp = fopen(filename, "rb");
  if (!fp) return WB_ERR_ARG;
  FileCloser guard(fp);
  if (!expect_magic(fp, magic)) return WB_ERR_ARG;
  int frames, fft_size, dims;
  int rc = read_matrix_header(fp, &frames, &fft_size, &dims);
  if (rc) return rc;
  for (int i = 0; i < frames; ++i) {
    double *row = rows ? rows[i] : contiguous + (size_t)i * row_stride;
    if (fread(row, 8, dims, fp) != (size_t)dims) return WB_ERR_ARG;
  }
  return WB_OK;
}

// ---- RIFF (tools/audioio.cpp) -----------------------------------------------------------------------
// The reference accepts exactly: "RIFF" <size> "WAVE" "fmt " 16 PCM(1) mono(1) fs <6 bytes> nbit, then skips
// forward to the first "data" tag (audioio.cpp:37-72, :89-112).
int wav_open(const char *filename, FILE **out, int *fs, int *nbit, int *length) {
  FILE *fp = fopen(filename, "rb");
  if (!fp) return WB_ERR_ARG;
  unsigned char h[36];
  bool ok = fread(h, 1, 36, fp) == 36 && !memcmp(h, "RIFF", 4) && !memcmp(h + 8, "WAVE", 4) && !memcmp(h + 12, "fmt ", 4) &&
            h[16] == 16 && !h[17] && !h[18] && !h[19] &&   // fmt chunk of 16 bytes
            h[20] == 1 && !h[21] &&                       // PCM
            h[22] == 1 && !h[23];                         // mono
  if (!ok) { fclose(fp); return WB_ERR_UNSUPPORTED; }
  *fs = (int)((uint32_t)h[24] | ((uint32_t)h[25] << 8) | ((uint32_t)h[26] << 16) | ((uint32_t)h[27] << 24));
  *nbit = h[34];                                          // low byte only, like the reference (audioio.cpp:97)
  if (*nbit < 8 || *nbit > 32 || (*nbit % 8)) { fclose(fp); return WB_ERR_UNSUPPORTED; }
  // scan for "data": a 'd' that does not start the tag is re-scanned from the next byte (audioio.cpp:100-108)
  bool found = false;
  int c;
  while ((c = fgetc(fp)) != EOF) {
    if (c != 'd') continue;
    unsigned char t[3];
    const size_t got = fread(t, 1, 3, fp);
    if (got == 3 && t[0] == 'a' && t[1] == 't' && t[2] == 'a') { found = true; break; }
    if (got != 3) break;
    fseek(fp, -3, SEEK_CUR);
  }
  unsigned char s[4];
  if (!found || fread(s, 1, 4, fp) != 4) { fclose(fp); return WB_ERR_UNSUPPORTED; }
  const int bytes = (int)((uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16) | ((uint32_t)s[3] << 24));
  *length = bytes / (*nbit / 8);
  *out = fp;
  return WB_OK;
}

// ---- sample-format kernels ------------------------------------------------------------------------
// 16-bit branch of wavread (audioio.cpp:232-249): (magnitude - sign_bias) / 2^15 == int16 / 32768 exactly
__global__ void pcm16_to_f64_kernel(const short *__restrict__ in, int n, double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)in[i] / 32768.0;
}
// wavwrite (audioio.cpp:176-180): truncation toward zero of x * 32767, clamped to the int16 range
__global__ void f64_to_pcm16_kernel(const double *__restrict__ in, int n, short *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int v = __double2int_rz(in[i] * 32767.0);   // saturating where the C cast would be undefined
    out[i] = (short)wb_max_i(-32768, wb_min_i(32767, v));
  }
}
__global__ void f64_to_f32_kernel(const double *__restrict__ in, size_t n, float *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = i; k < n; k += stride) out[k] = (float)in[k];   // round to nearest even
}

}  // namespace

int wb_pcm16_to_f64_run(const short *d_in, int n, double *d_out, cudaStream_t stream) {
  if (n <= 0) return WB_OK;
  WB_LAUNCH("pcm16_to_f64_kernel", pcm16_to_f64_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_in, n, d_out));
  WB_CUDA_CHECK(cudaGetLastError());
  return WB_OK;
}
int wb_f64_to_pcm16_run(const double *d_in, int n, short *d_out, cudaStream_t stream) {
  if (n <= 0) return WB_OK;
  WB_LAUNCH("f64_to_pcm16_kernel", f64_to_pcm16_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_in, n, d_out));
  WB_CUDA_CHECK(cudaGetLastError());
  return WB_OK;
}
int wb_f64_to_f32_run(const double *d_in, size_t n, float *d_out, cudaStream_t stream) {
  if (n == 0) return WB_OK;
  const size_t blocks = (n + 255) / 256;
  WB_LAUNCH("f64_to_f32_kernel", f64_to_f32_kernel<<<(unsigned)(blocks < (size_t)wb_sm_count() * 16 ? blocks : (size_t)wb_sm_count() * 16), 256, 0, stream>>>(d_in, n, d_out));
  WB_CUDA_CHECK(cudaGetLastError());
  return WB_OK;
}

extern "C" {

// ---- tools/parameterio.cpp ---------------------------------------------------------------------------
int wb_write_f0(const char *filename, int f0_length, double frame_period, const double *temporal_positions,
                const double *f0, int text_flag) {
  if (!filename || f0_length < 0 || (f0_length > 0 && !f0)) return WB_ERR_ARG;
  if (text_flag == 1) {   // parameterio.cpp:64-72: "%.5f %.5f\r\n" per frame
    if (f0_length > 0 && !temporal_positions) return WB_ERR_ARG;
    FILE *fp = fopen(filename, "w");
    if (!fp) return WB_ERR_ARG;
    FileCloser guard(fp);
    for (int i = 0; i < f0_length; ++i)
      if (fprintf(fp, "%.5f %.5f\r\n", temporal_positions[i], f0[i]) < 0) return WB_ERR_ARG;
    return WB_OK;
  }
  FILE *fp = fopen(filename, "wb");
  if (!fp) return WB_ERR_ARG;
  FileCloser guard(fp);
  const bool ok = fwrite("F0  ", 1, 4, fp) == 4 && put_tag_i32(fp, "NOF ", f0_length) && put_tag_f64(fp, "FP  ", frame_period) &&
                  fwrite(f0, 8, f0_length, fp) == (size_t)f0_length;
  return ok ? WB_OK : WB_ERR_ARG;
}

int wb_read_f0(const char *filename, double *temporal_positions, double *f0) {
  if (!filename || !f0) return WB_ERR_ARG;
  FILE *fp = fopen(filename, "rb");
  if (!fp) return WB_ERR_ARG;
  FileCloser guard(fp);
  if (!expect_magic(fp, "F0  ")) return WB_ERR_ARG;
  unsigned char h[20];
  if (fread(h, 1, 20, fp) != 20) return WB_ERR_ARG;
  int32_t frames;
  double frame_period;
  memcpy(&frames, h + 4, 4);
  memcpy(&frame_period, h + 12, 8);
  if (frames < 0 || fread(f0, 8, frames, fp) != (size_t)frames) return WB_ERR_ARG;
  if (temporal_positions)
    for (int i = 0; i < frames; ++i) temporal_positions[i] = i / 1000.0 * frame_period;   // parameterio.cpp:116-117
  return WB_OK;
}

// parameterio.cpp:121-147: looks at 13 consecutive 4-byte words from the start of the file for the tag
double wb_get_header_information(const char *filename, const char *parameter) {
  if (!filename || !parameter || strlen(parameter) != 4) return 0;
  FILE *fp = fopen(filename, "rb");
  if (!fp) return 0;
  FileCloser guard(fp);
  for (int i = 0; i < 13; ++i) {
    char word[4];
    if (fread(word, 1, 4, fp) != 4) return 0;
    if (memcmp(word, parameter, 4)) continue;
    if (!memcmp(parameter, "FP  ", 4)) {
      double v;
      return fread(&v, 8, 1, fp) == 1 ? v : 0;
    }
    int32_t v;
    return fread(&v, 4, 1, fp) == 1 ? (double)v : 0;
  }
  return 0;
}

int wb_write_spectral_envelope(const char *filename, int fs, int f0_length, double frame_period, int fft_size,
                               int number_of_dimensions, const double *const *spectrogram) {
  return write_matrix("SPEC", filename, fs, f0_length, frame_period, fft_size, number_of_dimensions, spectrogram, nullptr, 0);
}
int wb_read_spectral_envelope(const char *filename, double **spectrogram) {
  return read_matrix("SPEC", filename, spectrogram, nullptr, 0);
}
int wb_write_aperiodicity(const char *filename, int fs, int f0_length, double frame_period, int fft_size,
                          int number_of_dimensions, const double *const *aperiodicity) {
  return write_matrix("AP  ", filename, fs, f0_length, frame_period, fft_size, number_of_dimensions, aperiodicity, nullptr, 0);
}
int wb_read_aperiodicity(const char *filename, double **aperiodicity) {
  return read_matrix("AP  ", filename, aperiodicity, nullptr, 0);
}
int wb_write_parameter_matrix(int kind, const char *filename, int fs, int f0_length, double frame_period, int fft_size,
                              int number_of_dimensions, const double *matrix, long long row_stride) {
  if (kind != 0 && kind != 1) return WB_ERR_ARG;
  return write_matrix(kind == 0 ? "SPEC" : "AP  ", filename, fs, f0_length, frame_period, fft_size, number_of_dimensions,
                      nullptr, matrix, row_stride);
}
int wb_read_parameter_matrix(int kind, const char *filename, double *matrix, long long row_stride) {
  if (kind != 0 && kind != 1) return WB_ERR_ARG;
  return read_matrix(kind == 0 ? "SPEC" : "AP  ", filename, nullptr, matrix, row_stride);
}

// ---- tools/audioio.cpp -------------------------------------------------------------------------------
// audioio.cpp:121-184: always 16-bit mono PCM, whatever `nbit` says
int wb_wavwrite(const double *x, int x_length, int fs, int nbit, const char *filename) {
  (void)nbit;
  if (!filename || x_length < 0 || (x_length > 0 && !x)) return WB_ERR_ARG;
  FILE *fp = fopen(filename, "wb");
  if (!fp) return WB_ERR_ARG;
  FileCloser guard(fp);
  unsigned char h[44];
  auto u32 = [&](int at, uint32_t v) { h[at] = v & 255; h[at + 1] = (v >> 8) & 255; h[at + 2] = (v >> 16) & 255; h[at + 3] = (v >> 24) & 255; };
  auto u16 = [&](int at, uint32_t v) { h[at] = v & 255; h[at + 1] = (v >> 8) & 255; };
  memcpy(h, "RIFF", 4); u32(4, 36u + (uint32_t)x_length * 2u);
  memcpy(h + 8, "WAVEfmt ", 8); u32(16, 16); u16(20, 1); u16(22, 1);
  u32(24, (uint32_t)fs); u32(28, (uint32_t)fs * 2u); u16(32, 2); u16(34, 16);
  memcpy(h + 36, "data", 4); u32(40, (uint32_t)x_length * 2u);
  if (fwrite(h, 1, 44, fp) != 44) return WB_ERR_ARG;
  const int kChunk = 1 << 15;
  int16_t buf[kChunk];
  for (int a = 0; a < x_length; a += kChunk) {
    const int n = x_length - a < kChunk ? x_length - a : kChunk;
    for (int i = 0; i < n; ++i) {
      const double s = x[a + i] * 32767;
      const int v = s >= 2147483647.0 ? 2147483647 : (s <= -2147483648.0 ? (-2147483647 - 1) : (s != s ? 0 : static_cast<int>(s)));
      buf[i] = static_cast<int16_t>(v > 32767 ? 32767 : (v < -32768 ? -32768 : v));
    }
    if (fwrite(buf, 2, n, fp) != (size_t)n) return WB_ERR_ARG;
  }
  return WB_OK;
}

// audioio.cpp:186-226: 0 if the file cannot be opened, -1 on a header it does not accept
int wb_get_audio_length(const char *filename) {
  if (!filename) return 0;
  FILE *fp = nullptr;
  int fs, nbit, length;
  FILE *probe = fopen(filename, "rb");
  if (!probe) return 0;
  fclose(probe);
  if (wav_open(filename, &fp, &fs, &nbit, &length)) return -1;
  fclose(fp);
  return length;
}

// audioio.cpp:228-253 (+ :89-117): 8/16/24/32-bit little-endian signed PCM -> [-1, 1)
int wb_wavread(const char *filename, int *fs, int *nbit, double *x) {
  if (!filename || !fs || !nbit || !x) return WB_ERR_ARG;
  FILE *fp = nullptr;
  int length;
  int rc = wav_open(filename, &fp, fs, nbit, &length);
  if (rc) return rc;
  FileCloser guard(fp);
  const int qb = *nbit / 8;
  const double zero_line = pow(2.0, *nbit - 1);
  unsigned char s[4];
  for (int i = 0; i < length; ++i) {
    if (fread(s, 1, qb, fp) != (size_t)qb) return WB_ERR_ARG;
    double bias = 0.0, mag = 0.0;
    if (s[qb - 1] >= 128) { bias = zero_line; s[qb - 1] &= 0x7F; }
    for (int j = qb - 1; j >= 0; --j) mag = mag * 256.0 + s[j];
    x[i] = (mag - bias) / zero_line;
  }
  return WB_OK;
}

// raw 16-bit samples of a wav file (for the PCM entry points): `pcm` holds wb_get_audio_length() entries
int wb_wavread_pcm16(const char *filename, int *fs, short *pcm) {
  if (!filename || !fs || !pcm) return WB_ERR_ARG;
  FILE *fp = nullptr;
  int nbit, length;
  int rc = wav_open(filename, &fp, fs, &nbit, &length);
  if (rc) return rc;
  FileCloser guard(fp);
  if (nbit != 16) return WB_ERR_UNSUPPORTED;
  unsigned char s[2];
  for (int i = 0; i < length; ++i) {
    if (fread(s, 1, 2, fp) != 2) return WB_ERR_ARG;
    pcm[i] = (short)(uint16_t)(s[0] | (s[1] << 8));
  }
  return WB_OK;
}

}  // extern "C"
