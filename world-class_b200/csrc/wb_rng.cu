// randn() stream service: jump-ahead tables, fill and advance kernels.  See wb_rng.cuh.
#include "wb_rng.cuh"

#include <string.h>
#include <mutex>
#include <vector>

namespace {

struct Mat128 { uint32_t col[128][4]; };

// one randn() call as a state transition (outputs discarded)
void host_step_call(uint32_t s[4]) {
  uint32_t t[4] = {s[0], s[1], s[2], s[3]};
  (void)wb_randn_next(t);
  s[0] = t[0]; s[1] = t[1]; s[2] = t[2]; s[3] = t[3];
}

void mat_apply(const Mat128 &m, uint32_t s[4]) {
  uint32_t o[4] = {0, 0, 0, 0};
  for (int w = 0; w < 4; ++w)
    for (int b = 0; b < 32; ++b)
      if (s[w] >> b & 1u)
        for (int k = 0; k < 4; ++k) o[k] ^= m.col[w * 32 + b][k];
  memcpy(s, o, sizeof(o));
}

void mat_square(const Mat128 &m, Mat128 &out) {
  for (int c = 0; c < 128; ++c) {
    uint32_t v[4] = {m.col[c][0], m.col[c][1], m.col[c][2], m.col[c][3]};
    mat_apply(m, v);
    memcpy(out.col[c], v, sizeof(v));
  }
}

std::vector<Mat128> g_pow;          // host copies of P[b] = M^(2^b)
uint4 *g_d_pow = nullptr;           // device copies
WbRngState *g_d_state = nullptr;    // device-resident global stream state
std::once_flag g_once;
int g_init_status = WB_OK;

int do_init() {
  g_pow.resize(WB_RNG_NPOW);
  for (int c = 0; c < 128; ++c) {
    uint32_t e[4] = {0, 0, 0, 0};
    e[c >> 5] = 1u << (c & 31);
    host_step_call(e);
    memcpy(g_pow[0].col[c], e, sizeof(e));
  }
  for (int b = 1; b < WB_RNG_NPOW; ++b) mat_square(g_pow[b - 1], g_pow[b]);
  // 4-bit window tables (see wb_rng.cuh)
  std::vector<uint32_t> tab((size_t)WB_RNG_NPOW * WB_RNG_TAB_ENTRIES * 4);
  for (int b = 0; b < WB_RNG_NPOW; ++b)
    for (int n = 0; n < 32; ++n)
      for (int v = 0; v < 16; ++v) {
        uint32_t acc[4] = {0, 0, 0, 0};
        for (int bit = 0; bit < 4; ++bit)
          if (v >> bit & 1)
            for (int k = 0; k < 4; ++k) acc[k] ^= g_pow[b].col[n * 4 + bit][k];
        memcpy(&tab[(((size_t)b * 32 + n) * 16 + v) * 4], acc, sizeof(acc));
      }
  WB_CUDA_CHECK(cudaMalloc(&g_d_pow, tab.size() * sizeof(uint32_t)));
  WB_CUDA_CHECK(cudaMemcpy(g_d_pow, tab.data(), tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  WB_CUDA_CHECK(cudaMalloc(&g_d_state, sizeof(WbRngState)));
  // seed of the reference (world_matlabfunctions.cpp:244-247)
  const WbRngState s0 = {{123456789u, 362436069u, 521288629u, 88675123u}};
  WB_CUDA_CHECK(cudaMemcpy(g_d_state, &s0, sizeof(s0), cudaMemcpyHostToDevice));
  return WB_OK;
}

#define WB_RNG_CHUNK 64

// Two-level jump: warp 0 jumps cooperatively to the block's first stream position (one table
// load per lane and matrix), every thread then only jumps by tid * WB_RNG_CHUNK, which touches
// the same seven small tables in every block (L1-resident).
__global__ void __launch_bounds__(128) rng_fill_kernel(const WbRngState *__restrict__ state,
                                                       const uint4 *__restrict__ pow_tables,
                                                       const unsigned long long *__restrict__ d_count,
                                                       unsigned long long max_count, double *__restrict__ out) {
  __shared__ uint32_t s_base[4];
  const unsigned long long count = d_count ? min(*d_count, max_count) : max_count;
  const unsigned long long block_begin = (unsigned long long)blockIdx.x * blockDim.x * WB_RNG_CHUNK;
  if (block_begin >= count) return;
  if (threadIdx.x < 32) {
    uint32_t b[4] = {state->s[0], state->s[1], state->s[2], state->s[3]};
    wb_rng_jump_warp(pow_tables, b, block_begin);
    if (threadIdx.x == 0) { s_base[0] = b[0]; s_base[1] = b[1]; s_base[2] = b[2]; s_base[3] = b[3]; }
  }
  __syncthreads();
  const unsigned long long begin = block_begin + (unsigned long long)threadIdx.x * WB_RNG_CHUNK;
  if (begin >= count) return;
  uint32_t s[4] = {s_base[0], s_base[1], s_base[2], s_base[3]};
  wb_rng_jump(pow_tables, s, (unsigned long long)threadIdx.x * WB_RNG_CHUNK);
  const unsigned long long end = min(count, begin + WB_RNG_CHUNK);
  for (unsigned long long i = begin; i < end; ++i) out[i] = wb_randn_next(s);
}

__global__ void rng_advance_kernel(WbRngState *state, const uint4 *__restrict__ pow_tables,
                                   const unsigned long long *__restrict__ d_count) {
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    uint32_t s[4] = {state->s[0], state->s[1], state->s[2], state->s[3]};
    wb_rng_jump_warp(pow_tables, s, *d_count);
    __syncwarp();
    if (threadIdx.x == 0) { state->s[0] = s[0]; state->s[1] = s[1]; state->s[2] = s[2]; state->s[3] = s[3]; }
  }
}

}  // namespace

int wb_rng_init() {
  std::call_once(g_once, []() { g_init_status = do_init(); });
  return g_init_status;
}

const uint4 *wb_rng_tables() { return g_d_pow; }
WbRngState *wb_rng_global_state() { return g_d_state; }

void wb_rng_host_jump(uint32_t s[4], unsigned long long n) {
  for (int b = 0; n; ++b, n >>= 1)
    if (n & 1ull) mat_apply(g_pow[b], s);
}

int wb_rng_fill(const WbRngState *d_state, const unsigned long long *d_count_or_null,
                unsigned long long max_count, double *d_out, cudaStream_t stream) {
  if (max_count == 0) return WB_OK;
  const unsigned long long threads = (max_count + WB_RNG_CHUNK - 1) / WB_RNG_CHUNK;
  const int block = 128;
  const unsigned long long grid = (threads + block - 1) / block;
  WB_LAUNCH("rng_fill_kernel", rng_fill_kernel<<<(unsigned)grid, block, 0, stream>>>(d_state, g_d_pow, d_count_or_null, max_count, d_out));
  WB_CUDA_CHECK(cudaGetLastError());
  return WB_OK;
}

int wb_rng_advance(WbRngState *d_state, const unsigned long long *d_count, cudaStream_t stream) {
  WB_LAUNCH("rng_advance_kernel", rng_advance_kernel<<<1, 32, 0, stream>>>(d_state, g_d_pow, d_count));
  WB_CUDA_CHECK(cudaGetLastError());
  return WB_OK;
}
