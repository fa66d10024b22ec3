// Grid-wide exclusive scan of per-frame randn() call counts for LONG streams.
//
// The one-CTA scan of wb_common.cuh (wb_block_count_scan) is the right tool for an utterance (2001 frames:
// ~6 us), but it walks a whole-stream contour at ~2 us per 1024 frames -- 1.5 to 2.3 ms per call for one hour of
// audio, repeated by every shard of every rank (a fifth of the one-hour job).  Above WB_SCAN_SINGLE_CTA_MAX items
// the scan runs in three launches over the whole GPU: tile-local offsets + tile sums, a one-CTA scan of the tile
// sums, and the addition of the tile offsets.  The result is the same integer prefix sum.
#pragma once
#include "wb_common.cuh"
#include "wb_internal.h"

#define WB_SCAN_TILE_THREADS 256
#define WB_SCAN_TILE_ITEMS 8
#define WB_SCAN_TILE (WB_SCAN_TILE_THREADS * WB_SCAN_TILE_ITEMS)
#define WB_SCAN_SINGLE_CTA_MAX 16384

// count: a device functor `unsigned long long operator()(int i) const`.  Thread t of tile b owns items
// b TILE + t ITEMS .. + ITEMS - 1; offsets[i] <- exclusive prefix inside the tile, tile_sum[b] <- the tile's total.
template <typename F>
__global__ void __launch_bounds__(WB_SCAN_TILE_THREADS) wb_tile_count_kernel(F count, int n, unsigned long long *__restrict__ offsets,
                                                                             unsigned long long *__restrict__ tile_sum) {
  __shared__ unsigned long long s_warp[WB_SCAN_TILE_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long base = (long long)blockIdx.x * WB_SCAN_TILE + (long long)tid * WB_SCAN_TILE_ITEMS;
  unsigned long long c[WB_SCAN_TILE_ITEMS], s = 0;
#pragma unroll
  for (int q = 0; q < WB_SCAN_TILE_ITEMS; ++q) {
    const long long i = base + q;
    c[q] = i < n ? count((int)i) : 0ull;
    s += c[q];
  }
  unsigned long long incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  unsigned long long run = incl - s;
  for (int w = 0; w < warp; ++w) run += s_warp[w];
#pragma unroll
  for (int q = 0; q < WB_SCAN_TILE_ITEMS; ++q) {
    const long long i = base + q;
    if (i < n) offsets[i] = run;
    run += c[q];
  }
  if (tid == WB_SCAN_TILE_THREADS - 1) tile_sum[blockIdx.x] = run;
}

// the two shape-independent launches (wb_runtime.cu): offsets[n] <- total, *skip_mid <- *skip_in + *skip_add,
// *skip_out <- *skip_mid + total (each pointer optional), then offsets[i] += offset of i's tile
int wb_tile_scan_finish(unsigned long long *d_offsets, int n, unsigned long long *d_tile_sum, unsigned long long *d_tile_off,
                        int n_tiles, const unsigned long long *d_skip_in, const unsigned long long *d_skip_add,
                        unsigned long long *d_skip_mid, unsigned long long *d_skip_out, cudaStream_t stream);

// `scratch`: name of this call site's tile buffers on `ws` (sites that may run concurrently use different names)
template <typename F>
int wb_count_scan_tiles(F count, int n, unsigned long long *d_offsets, const unsigned long long *d_skip_in,
                        const unsigned long long *d_skip_add, unsigned long long *d_skip_mid,
                        unsigned long long *d_skip_out, WbWorkspace *ws, const char *scratch, cudaStream_t stream) {
  const int n_tiles = (n + WB_SCAN_TILE - 1) / WB_SCAN_TILE;
  unsigned long long *d_tiles = (unsigned long long *)ws->get(scratch, sizeof(unsigned long long) * (2 * (size_t)n_tiles + 2));
  if (!d_tiles) return WB_ERR_CUDA;
  WB_LAUNCH("count_tiles_kernel", wb_tile_count_kernel<F><<<n_tiles, WB_SCAN_TILE_THREADS, 0, stream>>>(count, n, d_offsets, d_tiles));
  WB_CUDA_CHECK(cudaGetLastError());
  return wb_tile_scan_finish(d_offsets, n, d_tiles, d_tiles + n_tiles, n_tiles, d_skip_in, d_skip_add, d_skip_mid, d_skip_out, stream);
}
