// Batched stand-alone transforms with the reference's world_fft semantics
// (/root/reference/include/world_fft.hpp:21-43, src/world_fft.cpp:31-77): used by the
// parity tests to pin the shared-memory FFT against the reference convention, and by
// callers that used fft_plan_dft_* + fft_execute directly.
#include "wb_internal.h"
#include "wb_fft.cuh"

namespace {

template <int LOG2N>
__global__ void r2c_kernel(const double *__restrict__ in, const cplx *__restrict__ T, cplx *__restrict__ out) {
  extern __shared__ double2 smem_raw[];
  cplx *S = smem_raw;
  double *W = reinterpret_cast<double *>(S);
  constexpr int n = 1 << LOG2N, NC = n / 2;
  const double *src = in + (size_t)blockIdx.x * n;
  cplx *dst = out + (size_t)blockIdx.x * (NC + 1);
  for (int j = threadIdx.x; j < n; j += blockDim.x) W[wb_didx(j)] = src[j];
  __syncthreads();
  wb_rfft_t<1, LOG2N - 1>(S, T, [&](int k, cplx X) { dst[k] = X; });
}

template <int LOG2N>
__global__ void c2r_kernel(const cplx *__restrict__ in, const cplx *__restrict__ T, double *__restrict__ out) {
  extern __shared__ double2 smem_raw[];
  cplx *S = smem_raw;
  double *W = reinterpret_cast<double *>(S);
  constexpr int n = 1 << LOG2N, NC = n / 2;
  const cplx *src = in + (size_t)blockIdx.x * (NC + 1);
  double *dst = out + (size_t)blockIdx.x * n;
  wb_irfft_t<-1, LOG2N - 1>(S, T, [&](int k) { return src[k]; });
  for (int j = threadIdx.x; j < n; j += blockDim.x) dst[j] = W[wb_didx(j)];
}

template <int SIGN, int LOG2N>
__global__ void c2c_kernel(const cplx *__restrict__ in, const cplx *__restrict__ T, cplx *__restrict__ out) {
  extern __shared__ double2 smem_raw[];
  cplx *S = smem_raw;
  constexpr int n = 1 << LOG2N, log2n = LOG2N;
  const cplx *src = in + (size_t)blockIdx.x * n;
  cplx *dst = out + (size_t)blockIdx.x * n;
  for (int j = threadIdx.x; j < n; j += blockDim.x) S[wb_sidx(j)] = src[j];
  __syncthreads();
  wb_cfft_dif_t<SIGN, LOG2N>(S, T);
  for (int k = threadIdx.x; k < n; k += blockDim.x) dst[k] = S[wb_sidx(wb_brev(k, log2n))];
}

int ilog2(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return ((1 << l) == n) ? l : -1;
}

}  // namespace

// kind: 0 = r2c (in: batch*n doubles, out: batch*(n/2+1) complex)
//       1 = c2r (in: batch*(n/2+1) complex, out: batch*n doubles)
//       2 = c2c forward, 3 = c2c backward (in/out: batch*n complex)
int wb_fft_batch_dev(int kind, const void *d_in, int n, int batch, void *d_out, cudaStream_t stream) {
  const int l = ilog2(n);
  if (l < 7 || n > 16384 || batch < 0) return WB_ERR_UNSUPPORTED;
  if (batch == 0) return WB_OK;
  const int threads = 256;
  int rc;
  if (kind == 0 || kind == 1) {
    const cplx *T = wb_twiddle_table(n);
    if (!T) return WB_ERR_CUDA;
    const size_t smem = sizeof(cplx) * wb_fft_slots(n / 2);
    if (kind == 0) {
      rc = WB_DISPATCH_LOG2(l, 7, 14, {
        if (cudaFuncSetAttribute(r2c_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WB_ERR_CUDA;
        WB_LAUNCH("r2c_kernel", r2c_kernel<L2><<<batch, threads, smem, stream>>>((const double *)d_in, T, (cplx *)d_out));
      });
    } else {
      rc = WB_DISPATCH_LOG2(l, 7, 14, {
        if (cudaFuncSetAttribute(c2r_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WB_ERR_CUDA;
        WB_LAUNCH("c2r_kernel", c2r_kernel<L2><<<batch, threads, smem, stream>>>((const cplx *)d_in, T, (double *)d_out));
      });
    }
  } else {
    if (n > 8192) return WB_ERR_UNSUPPORTED;
    const cplx *T = wb_twiddle_table(2 * n);
    if (!T) return WB_ERR_CUDA;
    const size_t smem = sizeof(cplx) * wb_fft_slots(n);
    if (kind == 2) {
      rc = WB_DISPATCH_LOG2(l, 7, 13, {
        if (cudaFuncSetAttribute(c2c_kernel<1, L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WB_ERR_CUDA;
        WB_LAUNCH("c2c_kernel", c2c_kernel<1, L2><<<batch, threads, smem, stream>>>((const cplx *)d_in, T, (cplx *)d_out));
      });
    } else {
      rc = WB_DISPATCH_LOG2(l, 7, 13, {
        if (cudaFuncSetAttribute(c2c_kernel<-1, L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WB_ERR_CUDA;
        WB_LAUNCH("c2c_kernel", c2c_kernel<-1, L2><<<batch, threads, smem, stream>>>((const cplx *)d_in, T, (cplx *)d_out));
      });
    }
  }
  if (rc) return rc;
  WB_CUDA_CHECK(cudaGetLastError());
  return WB_OK;
}
