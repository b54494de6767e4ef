// Device-wide exclusive prefix sum (hand-written, three-phase: tile reduce, scan of tile
// sums, tile down-sweep).  Used for order-preserving stream compaction in the
// post-processing path (contour starts in raster order, chain arena offsets, kept polygons).
#pragma once
#include "common.cuh"

namespace ocrb {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <class TOut>
__device__ __forceinline__ TOut block_exclusive_scan(TOut v, TOut *total, TOut *smem /* >= 32 */) {
  // returns exclusive prefix of v over the block (blockDim.x == SCAN_THREADS)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  TOut inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    TOut o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    TOut w = lane < (SCAN_THREADS / 32) ? smem[lane] : TOut(0);
    TOut winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      TOut o = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += o;
    }
    smem[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  TOut res = smem[warp] + inc - v;
  *total = smem[32];
  __syncthreads();
  return res;
}

struct ScanIdentity {
  template <class TIn, class TOut>
  static __device__ __forceinline__ TOut apply(TIn v) { return (TOut)v; }
};
struct ScanNonZero {
  template <class TIn, class TOut>
  static __device__ __forceinline__ TOut apply(TIn v) { return v != 0 ? TOut(1) : TOut(0); }
};

template <class TIn, class TOut, class Op>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_reduce_kernel(const TIn *__restrict__ in, int64_t n, TOut *__restrict__ tile_sums) {
  __shared__ TOut smem[33];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  TOut s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    int64_t i = base + (int64_t)j * SCAN_THREADS + threadIdx.x;
    if (i < n) s += Op::template apply<TIn, TOut>(in[i]);
  }
  TOut total;
  block_exclusive_scan<TOut>(s, &total, smem);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

template <class TIn, class TOut, class Op>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_downsweep_kernel(const TIn *__restrict__ in, int64_t n,
                                                                           const TOut *__restrict__ tile_offsets,
                                                                           TOut *__restrict__ out) {
  __shared__ TOut smem[33];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  TOut v[SCAN_ITEMS];
  TOut s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    int64_t i = base + j;
    v[j] = i < n ? Op::template apply<TIn, TOut>(in[i]) : TOut(0);
    s += v[j];
  }
  TOut total;
  TOut excl = block_exclusive_scan<TOut>(s, &total, smem) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    int64_t i = base + j;
    if (i < n) out[i] = excl;
    excl += v[j];
  }
}

// single-block scan for the (small) array of tile sums; also writes the grand total to
// out[n] so callers get offsets[n] == total.
template <class TOut>
__global__ void __launch_bounds__(SCAN_THREADS) scan_small_kernel(TOut *__restrict__ data, int64_t n, TOut *__restrict__ total_out) {
  __shared__ TOut smem[33];
  __shared__ TOut carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += SCAN_THREADS) {
    int64_t i = base + threadIdx.x;
    TOut v = i < n ? data[i] : TOut(0);
    TOut total;
    TOut e = block_exclusive_scan<TOut>(v, &total, smem);
    if (i < n) data[i] = e + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// exclusive scan of in[0..n) into out[0..n); out[n] receives the total.
// scratch must hold >= tiles + tiles/SCAN_TILE + 8 elements of TOut.
template <class TIn, class TOut, class Op = ScanIdentity>
int exclusive_scan(ocrb_ctx *ctx, const TIn *in, int64_t n, TOut *out, TOut *scratch) {
  if (n <= 0) {
    OCRB_CUDA(cudaMemsetAsync(out, 0, sizeof(TOut), ctx->stream));
    return OCRB_OK;
  }
  int64_t tiles = cdiv(n, SCAN_TILE);
  scan_tile_reduce_kernel<TIn, TOut, Op><<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(in, n, scratch);
  OCRB_TRY(check_launch(ctx, "scan_tile_reduce"));
  if (tiles <= 64 * SCAN_THREADS) {
    scan_small_kernel<TOut><<<1, SCAN_THREADS, 0, ctx->stream>>>(scratch, tiles, out + n);
    OCRB_TRY(check_launch(ctx, "scan_small"));
  } else {
    // two-level: scan the tile sums with the same machinery
    TOut *lvl2 = scratch + tiles + 1;
    OCRB_TRY((exclusive_scan<TOut, TOut, ScanIdentity>(ctx, scratch, tiles, scratch, lvl2)));
    // exclusive_scan wrote total at scratch[tiles]; move it to out[n]
    OCRB_CUDA(cudaMemcpyAsync(out + n, scratch + tiles, sizeof(TOut), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  scan_tile_downsweep_kernel<TIn, TOut, Op><<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(in, n, scratch, out);
  return check_launch(ctx, "scan_tile_downsweep");
}

inline size_t scan_scratch_elems(int64_t n) {
  int64_t tiles = cdiv(n > 0 ? n : 1, SCAN_TILE);
  return (size_t)(tiles + 1 + cdiv(tiles, SCAN_TILE) + 1 + 16);
}

}  // namespace ocrb
