// Exact reproduction of the reference's randn() stream on the GPU.
//
// Reference: /root/reference/src/world_matlabfunctions.cpp:243-264.  One randn() call is
// one partial shift (x<-y, y<-z, z<-w with w unchanged) followed by 12 xorshift128
// steps whose outputs (w >> 4) are summed; the value is sum / 2^28 - 6.  The state is a
// single process-global 128-bit word, so the value returned by the n-th call depends
// only on n.  xorshift128 is GF(2)-linear: the state after n calls is M^n s0 for a
// fixed 128x128 bit matrix M.  We precompute P[b] = M^(2^b) on the host and let every
// GPU thread jump straight to its own stream position.
#pragma once
#include "wb_common.cuh"

#define WB_RNG_NPOW 48  // jump distances up to 2^48 calls

struct WbRngState { uint32_t s[4]; };  // x, y, z, w

__host__ __device__ __forceinline__ double wb_randn_next(uint32_t (&s)[4]) {
  uint32_t x = s[0], y = s[1], z = s[2], w = s[3], t;
  x = y; y = z; z = w;  // partial shift: t is overwritten before use in the reference
  uint32_t tmp = 0;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    t = x ^ (x << 11);
    x = y; y = z; z = w;
    w = (w ^ (w >> 19)) ^ (t ^ (t >> 8));
    tmp += w >> 4;
  }
  s[0] = x; s[1] = y; s[2] = z; s[3] = w;
  return tmp / 268435456.0 - 6.0;
}

// One jump matrix is stored as a 4-bit window table: tab[n][v], n = 0..31 (nibble of the state),
// v = 0..15, holds the XOR of the matrix columns selected by v in nibble n.  A matrix-vector
// product is then 32 independent 16-byte loads (fully unrolled, no data-dependent loop).
#define WB_RNG_TAB_ENTRIES (32 * 16)
__device__ __forceinline__ void wb_rng_apply(const uint4 *__restrict__ tab, uint32_t (&s)[4]) {
  uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t nib = (s[w] >> (4 * k)) & 15u;
      const uint4 c = __ldg(&tab[(w * 8 + k) * 16 + nib]);
      o0 ^= c.x; o1 ^= c.y; o2 ^= c.z; o3 ^= c.w;
    }
  }
  s[0] = o0; s[1] = o1; s[2] = o2; s[3] = o3;
}

// same with the window table in shared memory
__device__ __forceinline__ void wb_rng_apply_shared(const uint4 *tab, uint32_t (&s)[4]) {
  uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t nib = (s[w] >> (4 * k)) & 15u;
      const uint4 c = tab[(w * 8 + k) * 16 + nib];
      o0 ^= c.x; o1 ^= c.y; o2 ^= c.z; o3 ^= c.w;
    }
  }
  s[0] = o0; s[1] = o1; s[2] = o2; s[3] = o3;
}

// s <- M^n s.  pow_tables: WB_RNG_NPOW window tables.
__device__ __forceinline__ void wb_rng_jump(const uint4 *__restrict__ pow_tables, uint32_t (&s)[4],
                                            unsigned long long n) {
  int b = 0;
  while (n) {
    if (n & 1ull) wb_rng_apply(pow_tables + (size_t)b * WB_RNG_TAB_ENTRIES, s);
    n >>= 1;
    ++b;
  }
}

// Warp-cooperative version: every lane passes the same state; lane l looks up nibble l and the
// partial results are XOR-reduced.  All 32 lanes must call.
__device__ __forceinline__ void wb_rng_jump_warp(const uint4 *__restrict__ pow_tables, uint32_t (&s)[4],
                                                 unsigned long long n) {
  const int lane = threadIdx.x & 31;
  int b = 0;
  while (n) {
    if (n & 1ull) {
      const uint32_t word = (lane >> 3) == 0 ? s[0] : ((lane >> 3) == 1 ? s[1] : ((lane >> 3) == 2 ? s[2] : s[3]));
      const uint32_t nib = (word >> (4 * (lane & 7))) & 15u;
      uint4 c = __ldg(&pow_tables[(size_t)b * WB_RNG_TAB_ENTRIES + lane * 16 + nib]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        c.x ^= __shfl_xor_sync(0xffffffffu, c.x, o);
        c.y ^= __shfl_xor_sync(0xffffffffu, c.y, o);
        c.z ^= __shfl_xor_sync(0xffffffffu, c.z, o);
        c.w ^= __shfl_xor_sync(0xffffffffu, c.w, o);
      }
      s[0] = c.x; s[1] = c.y; s[2] = c.z; s[3] = c.w;
    }
    n >>= 1;
    ++b;
  }
}

// ---- host API (wb_rng.cu) ---------------------------------------------------------------
// Device-resident global stream state (the reference's RNG is process-global).
int wb_rng_init();                         // builds tables, uploads, seeds the global state
const uint4 *wb_rng_tables();              // device pointer to the power tables
WbRngState *wb_rng_global_state();         // device pointer to the global state
void wb_rng_host_jump(uint32_t s[4], unsigned long long n);
// out[i] = value of the (*d_skip + i + 1)-th randn() call after state *d_state (d_skip null = 0); does not
// advance the state.
// d_skip2 (added to the skip) and d_count_sub (subtracted from the count) position a fill on a sub-range of a stage's
// draws straight from its offsets array: skip2 = count_sub = &offsets[begin], count = &offsets[end].
int wb_rng_fill(const WbRngState *d_state, const unsigned long long *d_skip_or_null,
                const unsigned long long *d_count_or_null, unsigned long long max_count, double *d_out,
                cudaStream_t stream, const unsigned long long *d_skip2 = nullptr,
                const unsigned long long *d_count_sub = nullptr);
// *d_state <- M^(*d_count + *d_count2) *d_state (either pointer may be null = 0)
int wb_rng_advance(WbRngState *d_state, const unsigned long long *d_count, const unsigned long long *d_count2_or_null,
                   cudaStream_t stream);

// warms L2 with the jump tables (asynchronous, see rng_prefetch_kernel)
int wb_rng_prefetch_tables(cudaStream_t stream);

// Where a stage's randn() draws sit in the stream and how it hands the position on.  The stage draws from
// *state advanced by *skip_in calls; once it has counted its own calls it writes skip_in + count to
// *skip_out.  Stand-alone stage calls set `advance` (the state itself is moved past the draws at the end);
// the whole-chain pipeline runs stages concurrently on several streams, chains them through the skip
// counters (ordered by the two events) and advances the state once at the end.
struct WbRngCursor {
  WbRngState *state = nullptr;
  const unsigned long long *skip_in = nullptr;   // device; null = 0
  unsigned long long *skip_out = nullptr;        // device; null = not needed
  bool advance = true;
  cudaEvent_t wait_skip_in = nullptr;            // waited for before *skip_in is first read
  cudaEvent_t record_skip_out = nullptr;         // recorded after *skip_out is written
};
