// Parameter modification between analysis and synthesis: F0 scaling and spectral stretching.
//
// Reference: the demo's ParameterModification, /root/reference/test/test.cpp:201-243 (SURVEY.md
// section 8f, N1), with interp1 / histc of /root/reference/src/world_matlabfunctions.cpp:136-182.
// One thread block per frame: the row's logarithm is staged in shared memory, every thread
// interpolates its output bins, exponentiates, and the bins above the stretched band edge repeat
// the last bin below it (ratio < 1).
#include "wb_internal.h"

namespace {

__global__ void f0_scale_kernel(const double *__restrict__ f0_in, int n, double shift, double *__restrict__ f0_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) f0_out[i] = f0_in[i] * shift;   // test.cpp:205-209
}

__global__ void __launch_bounds__(256) sp_stretch_kernel(double *__restrict__ sp, int fs, int fft_size, double ratio) {
  extern __shared__ double s_log[];  // bins
  const int bins = fft_size / 2 + 1;
  double *row = sp + (size_t)blockIdx.x * bins;
  for (int j = threadIdx.x; j < bins; j += blockDim.x) s_log[j] = log(row[j]);   // test.cpp:225-226
  __syncthreads();
  // knots x[k] = ratio * k / fft_size * fs (test.cpp:220), queries xi[j] = j / fft_size * fs (:221)
  auto knot = [&](int k) { return ratio * k / fft_size * fs; };
  const int cut = static_cast<int>(fft_size / 2.0 * ratio);   // test.cpp:232
  for (int j = threadIdx.x; j < bins; j += blockDim.x) {
    const double xi = static_cast<double>(j) / fft_size * fs;
    // histc: first knot strictly above xi, clamped to [1, bins - 1] (linear extrapolation beyond the knots)
    int k = static_cast<int>(j / ratio) + 1;
    k = wb_max_i(1, wb_min_i(bins - 1, k));
    while (k > 1 && knot(k - 1) > xi) --k;
    while (k < bins - 1 && !(knot(k) > xi)) ++k;
    const double x0 = knot(k - 1), x1 = knot(k);
    const double s = (xi - x0) / (x1 - x0);
    const double v = s_log[k - 1] + s * (s_log[k] - s_log[k - 1]);
    row[j] = exp(v);                                           // test.cpp:229-230
  }
  if (ratio >= 1.0) return;
  __syncthreads();
  // test.cpp:232-235: bins from `cut` on repeat the (already stretched) bin before them
  const double edge = row[cut - 1];
  __syncthreads();
  for (int j = cut + threadIdx.x; j < bins; j += blockDim.x) row[j] = edge;
}

}  // namespace

// d_f0_out may alias d_f0_in.  f0_shift is applied unless it is NaN; the spectrogram is stretched if ratio > 0.
int wb_parameter_modification_run(const double *d_f0_in, double *d_f0_out, int f0_length, double *d_sp, int fs,
                                  int fft_size, double f0_shift, double ratio, cudaStream_t stream) {
  if (f0_length <= 0) return WB_OK;
  if (f0_shift == f0_shift && d_f0_in && d_f0_out) {
    WB_LAUNCH("f0_scale_kernel", f0_scale_kernel<<<(f0_length + 255) / 256, 256, 0, stream>>>(d_f0_in, f0_length, f0_shift, d_f0_out));
  }
  if (ratio > 0.0 && d_sp) {
    const int bins = fft_size / 2 + 1;
    if (fft_size < 4 || static_cast<int>(fft_size / 2.0 * ratio) < 1) return WB_ERR_ARG;  // the reference indexes bin cut - 1
    const size_t smem = sizeof(double) * bins;
    if (smem > 200 * 1024) return WB_ERR_UNSUPPORTED;
    WB_CUDA_CHECK(cudaFuncSetAttribute(sp_stretch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WB_LAUNCH("sp_stretch_kernel", sp_stretch_kernel<<<f0_length, 256, smem, stream>>>(d_sp, fs, fft_size, ratio));
  }
  WB_CUDA_CHECK(cudaGetLastError());
  return WB_OK;
}
