// Coder / decoder of the spectral envelope (mel-cepstrum by DCT-via-FFT) and the aperiodicity
// (3 kHz band levels): one thread block per frame, BASELINE.json configs[4].
//
// Reference: /root/reference/src/codec.cpp
//   GetNumberOfAperiodicities :211-214, CodeAperiodicity :216-235, DecodeAperiodicity :237-265
//   (CheckVUV :30-40, GetAperiodicity :45-53), CodeSpectralEnvelope :267-296 (CodeOneFrame
//   :118-130, DCTForCodec :72-87, GetParametersForCoding :156-175), DecodeSpectralEnvelope
//   :298-325 (DecodeOneFrame :135-151, IDCTForCodec :92-113, GetParametersForDecoding :180-207).
//
// The interp1 resampling between the linear and the mel axis uses the same knots for every
// frame, so the segment indices and weights are tabulated once on the host with the reference's
// own expressions (glibc libm, like the reference).
#include "wb_internal.h"
#include "wb_fft.cuh"

#include <math.h>
#include <map>
#include <mutex>
#include <vector>

namespace {

const double kM0 = 1127.01048, kF0 = 700.0, kFloorFrequency = 40.0, kCeilFrequency = 20000.0;

inline double FrequencyToMel(double frequency) { return kM0 * log(frequency / kF0 + 1.0); }
inline double MelToFrequency(double mel) { return kF0 * (exp(mel / kM0) - 1.0); }

// histc + the `s` of interp1 (world_matlabfunctions.cpp:136-182) for sorted query points
void interp1_table(const std::vector<double> &x, const std::vector<double> &xi, std::vector<int> &k_out,
                   std::vector<double> &s_out) {
  const int x_length = (int)x.size(), xi_length = (int)xi.size();
  std::vector<int> k(xi_length, 0);
  {  // histc, transcribed statement by statement so that non-monotone knots behave alike
    int count = 1;
    int i = 0;
    for (; i < xi_length; ++i) {
      k[i] = 1;
      if (xi[i] >= x[0]) break;
    }
    for (; i < xi_length; ++i) {
      if (xi[i] < x[count]) {
        k[i] = count;
      } else {
        k[i--] = count++;
      }
      if (count == x_length) break;
    }
    count--;
    for (i++; i < xi_length; ++i) k[i] = count;
  }
  k_out = k;
  s_out.resize(xi_length);
  for (int i = 0; i < xi_length; ++i) {
    const double h = x[k[i]] - x[k[i] - 1];
    s_out[i] = (xi[i] - x[k[i] - 1]) / h;
  }
}

// ---- aperiodicity --------------------------------------------------------------------------
__global__ void code_ap_kernel(const double *__restrict__ ap, int f0_length, int fs, int fft_size, int n_ap,
                               double *__restrict__ coded) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= f0_length * n_ap) return;
  const int frame = g / n_ap, i = g % n_ap;
  const int bins = fft_size / 2 + 1;
  const double *row = ap + (size_t)frame * bins;
  // interp1Q(0, fs / fft_size, log_aperiodicity, bins, coarse_axis, ...) (codec.cpp:229-231)
  const double delta_x = static_cast<double>(fs) / fft_size;
  const double xi = WB_FREQ_INTERVAL * (i + 1.0);
  const int base = static_cast<int>((xi - 0) / delta_x);
  const double frac = (xi - 0) / delta_x - base;
  const double y0 = 20 * log10(row[base]);
  const double dy = (base >= bins - 1) ? 0.0 : 20 * log10(row[base + 1]) - y0;
  coded[(size_t)frame * n_ap + i] = y0 + dy * frac;
}

__global__ void decode_ap_kernel(const double *__restrict__ coded, int f0_length, int fs, int fft_size, int n_ap,
                                 double *__restrict__ ap) {
  const int frame = blockIdx.x;
  const int bins = fft_size / 2 + 1;
  double *out = ap + (size_t)frame * bins;
  const double *c = coded + (size_t)frame * n_ap;
  // CheckVUV (codec.cpp:30-40)
  double tmp = 0.0;
  for (int i = 0; i < n_ap; ++i) tmp += c[i];
  tmp /= n_ap;
  if (tmp > -0.5) {
    for (int i = threadIdx.x; i < bins; i += blockDim.x) out[i] = 1.0 - WB_SAFEGUARD;
    return;
  }
  const int nk = n_ap + 2;
  for (int i = threadIdx.x; i < bins; i += blockDim.x) {
    const double xi = static_cast<double>(fs) / fft_size * i;
    int k = 1;
    while (k < nk - 1) {
      if (xi < k * WB_FREQ_INTERVAL) break;
      ++k;
    }
    const double x0 = (k - 1) * WB_FREQ_INTERVAL;
    const double x1 = (k == nk - 1) ? fs / 2.0 : k * WB_FREQ_INTERVAL;
    const double y0 = (k - 1 == 0) ? -60.0 : c[k - 2];
    const double y1 = (k == nk - 1) ? -WB_SAFEGUARD : c[k - 1];
    const double s = (xi - x0) / (x1 - x0);
    const double v = y0 + s * (y1 - y0);
    out[i] = pow(10.0, v / 20.0);
  }
}

// ---- spectral envelope -----------------------------------------------------------------------
struct CodeSpParams {
  const double *sp; int f0_length; int fft_size; int nd;
  const int *k; const double *s;          // max_dimension entries: segment index / weight on the mel axis
  const cplx *weight;                     // max_dimension entries
  const cplx *tw;                         // max_dimension entries (real transform of size max_dimension)
  double *coded;                          // [f0_length][nd]
};

template <int LOG2M>  // max_dimension = fft_size / 2 = 2^LOG2M
__global__ void __launch_bounds__(256) code_sp_kernel(CodeSpParams p) {
  extern __shared__ double2 smem_raw[];
  constexpr int M = 1 << LOG2M;
  cplx *S = smem_raw;
  double *W = reinterpret_cast<double *>(S);
  double *logsp = reinterpret_cast<double *>(S + wb_fft_slots(M / 2));  // M + 1
  const int frame = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const double *row = p.sp + (size_t)frame * (M + 1);
  for (int j = tid; j <= M; j += nt) logsp[j] = log(row[j]);
  __syncthreads();
  // interp1 to the mel axis, then the even/odd permutation of DCTForCodec (codec.cpp:76-81)
  for (int j = tid; j < M; j += nt) {
    const int k = p.k[j];
    const double mel = logsp[k - 1] + p.s[j] * (logsp[k] - logsp[k - 1]);
    const int dst = (j & 1) ? (M - 1 - (j >> 1)) : (j >> 1);  // j = 2i -> i ; j = M - 2i - 1 -> i + M/2
    W[wb_didx(dst)] = mel;
  }
  __syncthreads();
  const double normalization = sqrt((double)M);
  double *out = p.coded + (size_t)frame * p.nd;
  const int nd = p.nd;
  wb_rfft_t<1, LOG2M - 1>(S, p.tw, [&](int k, cplx X) {
    if (k < nd) {
      const cplx w = p.weight[k];
      out[k] = (X.x * w.x - X.y * w.y) / normalization;
    }
  });
}

struct DecodeSpParams {
  const double *coded; int f0_length; int fft_size; int nd;
  const int *k; const double *s;          // bins entries: segment index / weight on the (max_dimension + 2)-knot axis
  const cplx *weight;                     // nd entries
  const cplx *tw;                         // 2 * max_dimension entries
  double *sp;                             // [f0_length][bins]
};

template <int LOG2M>
__global__ void __launch_bounds__(256) decode_sp_kernel(DecodeSpParams p) {
  extern __shared__ double2 smem_raw[];
  constexpr int M = 1 << LOG2M;
  cplx *S = smem_raw;
  double *ms = reinterpret_cast<double *>(S + wb_fft_slots(M));  // M + 2
  const int frame = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const double *c = p.coded + (size_t)frame * p.nd;
  const double normalization = sqrt((double)M);
  for (int i = tid; i < M; i += nt) {
    cplx z = make_double2(0.0, 0.0);
    if (i < p.nd) {
      const cplx w = p.weight[i];
      z.x = c[i] * w.x * normalization;
      z.y = -c[i] * w.y * normalization;
    }
    S[wb_sidx(i)] = z;
  }
  __syncthreads();
  wb_cfft_dif_t<-1, LOG2M>(S, p.tw);  // c2c FFT_BACKWARD (codec.cpp:106)
  // de-interleave into mel_spectrum[1..M] (codec.cpp:108-112, :142-144)
  for (int i = tid; i < M / 2; i += nt) {
    ms[1 + i * 2] = S[wb_sidx(wb_brev(i, LOG2M))].x;
    ms[1 + i * 2 + 1] = S[wb_sidx(wb_brev(M - i - 1, LOG2M))].x;
  }
  __syncthreads();
  if (tid == 0) { ms[0] = ms[1]; ms[M + 1] = ms[M]; }
  __syncthreads();
  double *out = p.sp + (size_t)frame * (M + 1);
  for (int j = tid; j <= M; j += nt) {
    const int k = p.k[j];
    const double v = ms[k - 1] + p.s[j] * (ms[k] - ms[k - 1]);
    out[j] = exp(v / M);
  }
}

int ilog2_exact(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return ((1 << l) == n) ? l : -1;
}

}  // namespace

int wb_code_aperiodicity_dev(const double *d_ap, int f0_length, int fs, int fft_size, double *d_coded,
                             cudaStream_t stream) {
  const int n_ap = wb_number_of_aperiodicities(fs);
  if (f0_length <= 0 || n_ap <= 0) return WB_OK;
  const int n = f0_length * n_ap;
  WB_LAUNCH("code_ap_kernel", code_ap_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_ap, f0_length, fs, fft_size, n_ap, d_coded));
  WB_CUDA_CHECK(cudaGetLastError());
  return WB_OK;
}

int wb_decode_aperiodicity_dev(const double *d_coded, int f0_length, int fs, int fft_size, double *d_ap,
                               cudaStream_t stream) {
  const int n_ap = wb_number_of_aperiodicities(fs);
  if (f0_length <= 0) return WB_OK;
  if (n_ap <= 0) return WB_ERR_UNSUPPORTED;
  WB_LAUNCH("decode_ap_kernel", decode_ap_kernel<<<f0_length, 256, 0, stream>>>(d_coded, f0_length, fs, fft_size, n_ap, d_ap));
  WB_CUDA_CHECK(cudaGetLastError());
  return WB_OK;
}

// Interpolation / weight tables of the spectral-envelope codec: functions of (fs, fft_size) -- and of the number
// of dimensions on the decoding side -- only.  Built once per key with the reference's host expressions
// (GetParametersForCoding codec.cpp:156-175, GetParametersForDecoding codec.cpp:180-207), uploaded once, never
// freed or rewritten: calls on any thread / stream share them read-only and a call is launches only (no
// synchronisation, no upload).
namespace {
struct CodecTables { int *k; double *s; cplx *w; };
struct CodecKey {
  bool decode; int fs, fft_size, nd;
  bool operator<(const CodecKey &o) const {
    if (decode != o.decode) return decode < o.decode;
    if (fs != o.fs) return fs < o.fs;
    if (fft_size != o.fft_size) return fft_size < o.fft_size;
    return nd < o.nd;
  }
};
std::mutex g_codec_mutex;
std::map<CodecKey, CodecTables> g_codec_tables;

const CodecTables *codec_tables(bool decode, int fs, int fft_size, int nd) {
  const CodecKey key = {decode, fs, fft_size, decode ? nd : 0};
  std::lock_guard<std::mutex> lock(g_codec_mutex);
  auto it = g_codec_tables.find(key);
  if (it != g_codec_tables.end()) return &it->second;
  const int M = fft_size / 2, bins = M + 1;
  const double floor_mel = FrequencyToMel(kFloorFrequency);
  const double ceil_f = (fs / 2.0 < kCeilFrequency) ? fs / 2.0 : kCeilFrequency;
  const double ceil_mel = FrequencyToMel(ceil_f);
  std::vector<int> k;
  std::vector<double> s;
  std::vector<cplx> weight;
  if (!decode) {
    std::vector<double> mel_axis(M), frequency_axis(M + 1, 0.0);  // frequency_axis[M] is never set (zero-filled, Q12)
    weight.resize(M);
    for (int i = 0; i < M; ++i) {
      mel_axis[i] = (ceil_mel - floor_mel) * i / M + floor_mel;
      weight[i].x = 2.0 * cos(i * WB_PI / fft_size) / sqrt((double)fft_size);
      weight[i].y = 2.0 * sin(i * WB_PI / fft_size) / sqrt((double)fft_size);
    }
    weight[0].x /= sqrt(2.0);
    for (int i = 0; i < M; ++i) frequency_axis[i] = FrequencyToMel(static_cast<double>(i) * fs / fft_size);
    interp1_table(frequency_axis, mel_axis, k, s);
  } else {
    weight.resize(nd);
    for (int i = 0; i < nd; ++i) {
      weight[i].x = cos(i * WB_PI / fft_size) * sqrt((double)fft_size);
      weight[i].y = sin(i * WB_PI / fft_size) * sqrt((double)fft_size);
    }
    weight[0].x /= sqrt(2.0);
    std::vector<double> mel_axis(M + 2), frequency_axis(bins);
    for (int i = 0; i < M; ++i) mel_axis[i + 1] = MelToFrequency((ceil_mel - floor_mel) * i / M + floor_mel);
    mel_axis[0] = 0;
    mel_axis[M + 1] = fs / 2.0;
    for (int i = 0; i < bins; ++i) frequency_axis[i] = static_cast<double>(i) * fs / fft_size;
    interp1_table(mel_axis, frequency_axis, k, s);
  }
  CodecTables t = {nullptr, nullptr, nullptr};
  if (cudaMalloc(&t.k, sizeof(int) * k.size()) != cudaSuccess || cudaMalloc(&t.s, sizeof(double) * s.size()) != cudaSuccess ||
      cudaMalloc(&t.w, sizeof(cplx) * weight.size()) != cudaSuccess)
    return nullptr;
  if (cudaMemcpy(t.k, k.data(), sizeof(int) * k.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(t.s, s.data(), sizeof(double) * s.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(t.w, weight.data(), sizeof(cplx) * weight.size(), cudaMemcpyHostToDevice) != cudaSuccess)
    return nullptr;
  return &(g_codec_tables[key] = t);
}
}  // namespace

int wb_code_spectral_envelope_dev(WbWorkspace *ws, const double *d_sp, int f0_length, int fs, int fft_size, int nd,
                                  double *d_coded, cudaStream_t stream) {
  if (f0_length <= 0) return WB_OK;
  const int M = fft_size / 2, l = ilog2_exact(M);
  if (l < 7 || l > 12 || nd < 1 || nd > M) return WB_ERR_UNSUPPORTED;
  const CodecTables *tab = codec_tables(false, fs, fft_size, nd);
  if (!tab) return WB_ERR_CUDA;
  int *d_k = tab->k;
  double *d_s = tab->s;
  cplx *d_w = tab->w;
  (void)ws;
  CodeSpParams p;
  p.sp = d_sp; p.f0_length = f0_length; p.fft_size = fft_size; p.nd = nd; p.k = d_k; p.s = d_s; p.weight = d_w;
  p.tw = wb_twiddle_table(M);
  p.coded = d_coded;
  if (!p.tw) return WB_ERR_CUDA;
  const size_t smem = sizeof(cplx) * wb_fft_slots(M / 2) + sizeof(double) * (M + 2);
  int rc = WB_DISPATCH_LOG2(l, 7, 12, {
    if (cudaFuncSetAttribute(code_sp_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WB_ERR_CUDA;
    WB_LAUNCH("code_sp_kernel", code_sp_kernel<L2><<<f0_length, 256, smem, stream>>>(p));
  });
  if (rc) return rc;
  WB_CUDA_CHECK(cudaGetLastError());
  return WB_OK;
}

int wb_decode_spectral_envelope_dev(WbWorkspace *ws, const double *d_coded, int f0_length, int fs, int fft_size,
                                    int nd, double *d_sp, cudaStream_t stream) {
  if (f0_length <= 0) return WB_OK;
  const int M = fft_size / 2, l = ilog2_exact(M), bins = M + 1;
  if (l < 7 || l > 12 || nd < 1 || nd > M) return WB_ERR_UNSUPPORTED;
  const CodecTables *tab = codec_tables(true, fs, fft_size, nd);
  if (!tab) return WB_ERR_CUDA;
  int *d_k = tab->k;
  double *d_s = tab->s;
  cplx *d_w = tab->w;
  (void)ws; (void)bins;
  DecodeSpParams p;
  p.coded = d_coded; p.f0_length = f0_length; p.fft_size = fft_size; p.nd = nd; p.k = d_k; p.s = d_s; p.weight = d_w;
  p.tw = wb_twiddle_table(2 * M);
  p.sp = d_sp;
  if (!p.tw) return WB_ERR_CUDA;
  const size_t smem = sizeof(cplx) * wb_fft_slots(M) + sizeof(double) * (M + 4);
  int rc = WB_DISPATCH_LOG2(l, 7, 12, {
    if (cudaFuncSetAttribute(decode_sp_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WB_ERR_CUDA;
    WB_LAUNCH("decode_sp_kernel", decode_sp_kernel<L2><<<f0_length, 256, smem, stream>>>(p));
  });
  if (rc) return rc;
  WB_CUDA_CHECK(cudaGetLastError());
  return WB_OK;
}
