// Connected-component labelling of a binary map: ONE union-find over both pixel classes
// (foreground 8-connected, background 4-connected), block-local in shared memory with
// warp-level run merges, followed by a global seam-merge pass and a flatten pass.
//
// Replaces the sequential raster scan inside imageproc::contours::find_contours
// (called at metrics.rs:78-81): after flattening, label[i] is the raster-first pixel index
// of i's component, which is exactly where Suzuki–Abe starts that component's outer border
// (foreground) or — one pixel to the west — its hole border (background); SURVEY A.1.
//
// HBM traffic: reads the u8 bitmap (1 B/px) and writes i32 labels (4 B/px) in the local
// pass — except for tiles without a foreground pixel, where only the tile origin's label is
// written (ccl.cuh: ccl_parent answers for the rest); the seam pass touches border labels again.
#include "ccl.cuh"
#include "common.cuh"
#include "scan.cuh"

namespace ocrb {

__device__ __forceinline__ int uf_find(const int *L, int a) {
  int p = L[a];
  while (p != a) {
    a = p;
    p = L[a];
  }
  return a;
}

__device__ __forceinline__ int uf_find_volatile(volatile int *L, int a) {
  int p = L[a];
  while (p != a) {
    a = p;
    p = L[a];
  }
  return a;
}

// lock-free union keeping the smaller index as root
__device__ __forceinline__ void uf_union(int *L, int a, int b) {
  for (;;) {
    a = uf_find_volatile(L, a);
    b = uf_find_volatile(L, b);
    if (a == b) return;
    if (a > b) { int t = a; a = b; b = t; }
    int old = atomicMin(&L[b], a);
    if (old == b) return;
    b = old;
  }
}

// ---------------------------------------------------------------------------------------
// pass 1: tile-local.  256 threads = 8 warps; warp w owns tile rows 4w .. 4w+3.
// Warp-level merge: every pixel starts labelled with the first pixel of its horizontal run
// (ballot + bit scan, no atomics).  Vertical / diagonal links are then united in shared
// memory, one union per pair of overlapping runs.  Output: parent = image-local linear index
// of the tile-local root.
// ---------------------------------------------------------------------------------------
constexpr int CCL_THREADS = 256;
constexpr int CCL_ROWS_PER_WARP = CCL_TH / (CCL_THREADS / 32);
constexpr int CCL_STRIP = 16;  // tiles per CTA: a 32-row x 512-column strip (16 KB, four 16-byte loads per thread)

// one tile with foreground: union-find in shared memory (L: CCL_TW * CCL_TH ints)
__device__ __forceinline__ void ccl_tile(const uint8_t *__restrict__ bm, int H, int W, int tx, int ty, int *__restrict__ labels_img, int *L,
                                         uint32_t *rowbits, uint32_t *rowvalid) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = tx * CCL_TW + lane;
  int fgv[CCL_ROWS_PER_WARP];
#pragma unroll
  for (int k = 0; k < CCL_ROWS_PER_WARP; ++k) {  // all loads first (the strip pass just touched these lines: L1 hits)
    const int y = ty * CCL_TH + warp * CCL_ROWS_PER_WARP + k;
    fgv[k] = (x < W && y < H) ? (bm[(int64_t)y * W + x] != 0) : 0;
  }
#pragma unroll
  for (int k = 0; k < CCL_ROWS_PER_WARP; ++k) {
    const int r = warp * CCL_ROWS_PER_WARP + k, y = ty * CCL_TH + r;
    const uint32_t valid = __ballot_sync(0xffffffffu, x < W && y < H);
    const uint32_t bits = __ballot_sync(0xffffffffu, fgv[k]);
    // run start of this lane within its row: pixels of the same class contiguous to the left
    const uint32_t same = fgv[k] ? bits : (~bits & valid);
    const uint32_t below = (~same) & ((1u << lane) - 1u);  // lanes to the left that break the run
    const int run_start = below ? (32 - __clz(below)) : 0;
    L[r * CCL_TW + lane] = r * CCL_TW + run_start;
    if (lane == 0) { rowbits[r] = bits; rowvalid[r] = valid; }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < CCL_ROWS_PER_WARP; ++k) {
    const int r = warp * CCL_ROWS_PER_WARP + k;
    const uint32_t valid = rowvalid[r], bits = rowbits[r];
    if (r == 0 || !((valid >> lane) & 1)) continue;
    // One union per pair of overlapping runs, not per pixel: a vertical link is redundant when
    // the pixel to the left is in my run and the pixel above it is in the run above me (the
    // leftmost pixel of the overlap makes the link).  Diagonal links (8-connectivity of the
    // foreground) are only needed from the ends of a run.
    const int fg = (bits >> lane) & 1;
    const int self = r * CCL_TW + lane;
    const uint32_t same = fg ? bits : (~bits & valid);
    const uint32_t up = rowbits[r - 1];
    const uint32_t up_same = fg ? up : (~up & valid);  // pixels above of my class
    if ((up_same >> lane) & 1) {
      const bool left_mine = lane > 0 && ((same >> (lane - 1)) & 1);
      const bool upleft_same = lane > 0 && ((up_same >> (lane - 1)) & 1);
      if (!(left_mine && upleft_same)) uf_union(L, self, self - CCL_TW);
    } else if (fg) {
      const bool left_fg = lane > 0 && ((bits >> (lane - 1)) & 1);
      const bool right_fg = lane < 31 && ((bits >> (lane + 1)) & 1);
      if (lane > 0 && !left_fg && ((up >> (lane - 1)) & 1)) uf_union(L, self, self - CCL_TW - 1);
      if (lane < 31 && !right_fg && ((up >> (lane + 1)) & 1)) uf_union(L, self, self - CCL_TW + 1);
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < CCL_ROWS_PER_WARP; ++k) {
    const int r = warp * CCL_ROWS_PER_WARP + k, y = ty * CCL_TH + r;
    if (x < W && y < H) {
      const int root = uf_find(L, r * CCL_TW + lane);
      const int rx = tx * CCL_TW + (root & 31), ry = ty * CCL_TH + (root >> 5);
      labels_img[(int64_t)y * W + x] = ry * W + rx;
    }
  }
  __syncthreads();  // L / rowbits are reused by the strip's next tile
}

// One CTA per strip of CCL_STRIP tiles.  The strip is read once with 16-byte loads (thread = row, 16-pixel chunk) to find
// the tiles that hold foreground at all: on a document page most do not, and such a tile is one 4-connected background
// rectangle — only its origin's label is stored (ccl.cuh).  Tiles with foreground run the union-find one after another.
__global__ void __launch_bounds__(CCL_THREADS) ccl_local_kernel(const uint8_t *__restrict__ bitmap, int H, int W,
                                                                 int tiles_x, int tiles_y, int strips_x, int *__restrict__ labels,
                                                                 uint8_t *__restrict__ tile_empty) {
  __shared__ int L[CCL_TW * CCL_TH];
  __shared__ uint32_t rowbits[CCL_TH], rowvalid[CCL_TH];
  __shared__ int s_any[CCL_STRIP];
  const int strips = strips_x * tiles_y;
  const int strip = blockIdx.x % strips, b = blockIdx.x / strips;
  const int sx = strip % strips_x, ty = strip / strips_x;
  const uint8_t *bm = bitmap + (int64_t)b * H * W;
  int *labels_img = labels + (int64_t)b * H * W;
  if (threadIdx.x < CCL_STRIP) s_any[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int it = 0; it < CCL_TH * CCL_STRIP * 2 / CCL_THREADS; ++it) {
    const int idx = threadIdx.x + it * CCL_THREADS;
    const int r = idx / (CCL_STRIP * 2), c = idx - r * (CCL_STRIP * 2);  // 32 rows x (2 * CCL_STRIP) chunks of 16 pixels
    const int y = ty * CCL_TH + r, x = sx * CCL_STRIP * CCL_TW + c * 16;
    bool nz = false;
    if (y < H && x < W) {
      const uint8_t *p = bm + (int64_t)y * W + x;
      if (x + 16 <= W && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        const uint4 v = *reinterpret_cast<const uint4 *>(p);
        nz = (v.x | v.y | v.z | v.w) != 0;
      } else {
        for (int e = 0; e < 16 && x + e < W; ++e) nz |= p[e] != 0;
      }
    }
    if (nz) s_any[c >> 1] = 1;  // benign race: every writer stores 1
  }
  __syncthreads();
  if (threadIdx.x < CCL_STRIP) {  // one thread per tile of the strip: the flag, and the only label a tile without foreground stores
    const int tx = sx * CCL_STRIP + threadIdx.x;
    if (tx < tiles_x) {
      const int any_fg = s_any[threadIdx.x];
      tile_empty[(int64_t)b * tiles_x * tiles_y + ty * tiles_x + tx] = (uint8_t)!any_fg;  // the seam pass and the label consumers read this
      if (!any_fg) {
        const int origin = ty * CCL_TH * W + tx * CCL_TW;
        labels_img[origin] = origin;
      }
    }
  }
#pragma unroll 1
  for (int t = 0; t < CCL_STRIP; ++t) {
    const int tx = sx * CCL_STRIP + t;
    if (tx >= tiles_x) break;
    if (s_any[t]) ccl_tile(bm, H, W, tx, ty, labels_img, L, rowbits, rowvalid);
  }
}

// ---------------------------------------------------------------------------------------
// pass 2: seams.  A pixel whose W / NW / N / NE neighbour lies in another tile unites with
// it in global memory (same adjacency rules as pass 1).
// ---------------------------------------------------------------------------------------
// The lanes of a warp walk a tile's border pixels (top row, left column, right column).  A tile without foreground whose
// left and upper neighbours have none either needs exactly one link (its corner pixel) — on a document page that is almost
// every tile — so a CTA first gives every one of its 256 tiles ONE thread, which does that single link itself, and only
// the remaining tiles are walked by whole warps.
// border pixels a tile has to look at (k < k_end of top row | left column | right column): between tiles without
// foreground the corner pixel (k = 0, which always runs the full rules) makes the one link needed:
//   * right column: only foreground pixels link diagonally from there;
//   * left column below the corner: the link to the left tile is redundant when that tile is empty too;
//   * top row right of the corner: the link upwards is skipped by the run rule when the upper tile is empty too.
struct SeamTile { int b, tx, ty, k_end; bool empty, left_empty, up_empty; };
__device__ __forceinline__ SeamTile seam_tile_of(int64_t tile, int tiles_x, int tiles_y, const uint8_t *__restrict__ tile_empty) {
  constexpr int PER_TILE = CCL_TW + 2 * CCL_TH;
  SeamTile t;
  t.b = (int)(tile / ((int64_t)tiles_x * tiles_y));
  const int tt = (int)(tile % ((int64_t)tiles_x * tiles_y));
  t.tx = tt % tiles_x; t.ty = tt / tiles_x;
  const uint8_t *te = tile_empty + (int64_t)t.b * tiles_x * tiles_y;
  t.empty = te[tt] != 0;
  t.left_empty = t.tx > 0 && te[tt - 1]; t.up_empty = t.ty > 0 && te[tt - tiles_x];
  t.k_end = !t.empty ? PER_TILE : (((t.tx > 0 && !t.left_empty) ? CCL_TW + CCL_TH : ((t.ty > 0 && !t.up_empty) ? CCL_TW : 1)));
  return t;
}
__device__ __forceinline__ void seam_border_pixel(int k, const SeamTile &t, const uint8_t *__restrict__ bitmap, int H, int W, int tiles_x, int tiles_y,
                                                  int *__restrict__ labels, const uint8_t *__restrict__ tile_empty) {
  const int64_t HW = (int64_t)H * W;
  const uint8_t *te = tile_empty + (int64_t)t.b * tiles_x * tiles_y;
  const uint8_t *bm = bitmap + t.b * HW;
  int *L = labels + t.b * HW;
  auto node = [&](int px, int py) {  // a pixel of a tile without foreground stands for its tile origin (ccl.cuh)
    const int ttx = px / CCL_TW, tty = py / CCL_TH;
    return te[tty * tiles_x + ttx] ? tty * CCL_TH * W + ttx * CCL_TW : py * W + px;
  };
  int x, y;
  if (k < CCL_TW) { x = t.tx * CCL_TW + k; y = t.ty * CCL_TH; }
  else if (k < CCL_TW + CCL_TH) { x = t.tx * CCL_TW; y = t.ty * CCL_TH + (k - CCL_TW); }
  else { x = t.tx * CCL_TW + CCL_TW - 1; y = t.ty * CCL_TH + (k - CCL_TW - CCL_TH); }
  if (x >= W || y >= H) return;
  const bool on_left = (x % CCL_TW) == 0, on_right = (x % CCL_TW) == CCL_TW - 1, on_top = (y % CCL_TH) == 0;
  // corner pixels appear in two of the three groups: let the top-row instance do the work
  if (k >= CCL_TW && on_top) return;
  if (t.empty) {
    if (k >= CCL_TW && t.left_empty) return;
    if (k > 0 && k < CCL_TW && t.up_empty) return;
  }
  const int i = y * W + x;
  const int fg = bm[i] != 0;
  const int self = node(x, y);
  if (on_left && x > 0 && (bm[i - 1] != 0) == fg) uf_union(L, self, node(x - 1, y));
  if (y > 0) {
    const int n_fg = bm[i - W] != 0;
    if (on_top && n_fg == fg) {
      // same redundancy rule as the tile-local pass, along the whole image row
      const bool skip = x > 0 && (bm[i - 1] != 0) == fg && (bm[i - W - 1] != 0) == fg;
      if (!skip) uf_union(L, self, node(x, y - 1));
    }
    if (fg && !n_fg) {  // foreground pixels are never in an "empty" tile
      if (x > 0 && (on_top || on_left) && bm[i - W - 1] != 0) uf_union(L, i, i - W - 1);
      if (x + 1 < W && (on_top || on_right) && bm[i - W + 1] != 0) uf_union(L, i, i - W + 1);
    }
  }
}

// first kernel: one THREAD per tile.  A tile that only needs its corner pixel (on a document page almost every tile) is
// finished by that thread; the others are appended to a list (warp-aggregated atomics) ...
__global__ void __launch_bounds__(256) ccl_seam_kernel(const uint8_t *__restrict__ bitmap, int H, int W, int B, int tiles_x, int tiles_y,
                                                       int *__restrict__ labels, const uint8_t *__restrict__ tile_empty, int *__restrict__ list) {
  const int64_t tile = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t tiles = (int64_t)tiles_x * tiles_y * B;
  bool heavy = false;
  if (tile < tiles) {
    const SeamTile t = seam_tile_of(tile, tiles_x, tiles_y, tile_empty);
    if (t.k_end == 1) seam_border_pixel(0, t, bitmap, H, W, tiles_x, tiles_y, labels, tile_empty);
    else heavy = true;
  }
  const uint32_t m = __ballot_sync(0xffffffffu, heavy);
  if (m) {
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(&list[0], __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (heavy) list[1 + base + __popc(m & ((1u << lane) - 1))] = (int)tile;
  }
}
// ... which the second kernel walks with one WARP per listed tile (grid-stride: the launch does not know the count)
__global__ void __launch_bounds__(256) ccl_seam_heavy_kernel(const uint8_t *__restrict__ bitmap, int H, int W, int tiles_x, int tiles_y,
                                                             int *__restrict__ labels, const uint8_t *__restrict__ tile_empty,
                                                             const int *__restrict__ list) {
  const int lane = threadIdx.x & 31;
  const int n = list[0];
  const int warps = (int)((gridDim.x * blockDim.x) >> 5);
  for (int j = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); j < n; j += warps) {
    const SeamTile t = seam_tile_of(list[1 + j], tiles_x, tiles_y, tile_empty);
    for (int k = lane; k < t.k_end; k += 32) seam_border_pixel(k, t, bitmap, H, W, tiles_x, tiles_y, labels, tile_empty);
  }
}

// pass 3: flatten (every pixel points at its root = raster-first pixel of its component)
__global__ void ccl_flatten_kernel(int H, int W, int B, int *__restrict__ labels, CclTiles tiles) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t HW = (int64_t)H * W;
  if (idx >= HW * B) return;
  int b = (int)(idx / HW);
  int i = (int)(idx % HW);
  int *L = labels + b * HW;
  // roots are final here and only non-root entries are rewritten, so concurrent walks see consistent chains
  const int root = ccl_find_px(L, tiles, b, i, W);
  L[i] = root;
}

// flatten = false leaves a forest whose roots are final (label[i] == i  <=>  i is the raster-first
// pixel of its component); consumers that need the root of an arbitrary pixel call ccl_find.
int launch_ccl(ocrb_ctx *ctx, const uint8_t *bitmap, int B, int H, int W, int *labels, bool flatten) {
  int tiles_x = (int)cdiv(W, CCL_TW), tiles_y = (int)cdiv(H, CCL_TH);
  int64_t blocks = (int64_t)tiles_x * tiles_y * B;
  OCRB_TRY(ctx->ccl_tile_empty.reserve((size_t)blocks));
  uint8_t *tile_empty = ctx->ccl_tile_empty.as<uint8_t>();
  const int strips_x = (int)cdiv(tiles_x, CCL_STRIP);
  ccl_local_kernel<<<(unsigned)((int64_t)strips_x * tiles_y * B), CCL_THREADS, 0, ctx->stream>>>(bitmap, H, W, tiles_x, tiles_y, strips_x, labels, tile_empty);
  OCRB_TRY(check_launch(ctx, "ccl_local"));
  int64_t n = (int64_t)B * H * W;
  OCRB_REQUIRE(blocks < ((int64_t)1 << 31), "ccl: too many tiles");
  OCRB_TRY(ctx->ccl_seam_list.reserve((size_t)(blocks + 1) * 4));
  int *seam_list = ctx->ccl_seam_list.as<int>();
  OCRB_CUDA(cudaMemsetAsync(seam_list, 0, 4, ctx->stream));
  ccl_seam_kernel<<<(unsigned)cdiv(blocks, 256), 256, 0, ctx->stream>>>(bitmap, H, W, B, tiles_x, tiles_y, labels, tile_empty, seam_list);
  OCRB_TRY(check_launch(ctx, "ccl_seam"));
  {
    const int64_t want = cdiv(blocks * 32, 256), cap = (int64_t)ctx->sm_count * 8;
    ccl_seam_heavy_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, ctx->stream>>>(bitmap, H, W, tiles_x, tiles_y, labels, tile_empty, seam_list);
    OCRB_TRY(check_launch(ctx, "ccl_seam_heavy"));
  }
  if (!flatten) return OCRB_OK;
  CclTiles tiles = {tile_empty, tiles_x, tiles_x * tiles_y};
  ccl_flatten_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(H, W, B, labels, tiles);
  return check_launch(ctx, "ccl_flatten");
}

// ---------------------------------------------------------------------------------------
// test hook: canonical numbering of the foreground components (1..n in raster order of
// their first pixel, 0 = background) — what scipy.ndimage.label(structure=ones(3,3)) gives.
// ---------------------------------------------------------------------------------------
__global__ void ccl_fg_root_flag_kernel(const uint8_t *__restrict__ bitmap, const int *__restrict__ labels, int64_t HW,
                                        int64_t n, uint8_t *__restrict__ flag) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  int i = (int)(idx % HW);
  flag[idx] = (bitmap[idx] != 0 && labels[idx] == i) ? 1 : 0;
}

__global__ void ccl_canonical_kernel(const uint8_t *__restrict__ bitmap, const int *__restrict__ labels,
                                     const int *__restrict__ rank, int64_t HW, int64_t n, int *__restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  int64_t b = idx / HW;
  if (bitmap[idx] == 0) {
    out[idx] = 0;
    return;
  }
  int64_t root = b * HW + labels[idx];
  out[idx] = rank[root] - rank[b * HW] + 1;
}

int ccl_canonical_labels(ocrb_ctx *ctx, const uint8_t *bitmap_dev, int B, int H, int W, int *labels_out_dev,
                         int *n_components_host) {
  int64_t n = (int64_t)B * H * W, HW = (int64_t)H * W;
  DevBuf lab, flag, rank, scratch;
  int rc = OCRB_OK;
  do {
    if ((rc = lab.reserve(n * 4)) || (rc = flag.reserve(n)) || (rc = rank.reserve((n + 1) * 4)) ||
        (rc = scratch.reserve(scan_scratch_elems(n) * 4)))
      break;
    if ((rc = launch_ccl(ctx, bitmap_dev, B, H, W, lab.as<int>(), true))) break;
    ccl_fg_root_flag_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(bitmap_dev, lab.as<int>(), HW, n, flag.as<uint8_t>());
    if ((rc = check_launch(ctx, "ccl_fg_root_flag"))) break;
    if ((rc = exclusive_scan<uint8_t, int>(ctx, flag.as<uint8_t>(), n, rank.as<int>(), scratch.as<int>()))) break;
    ccl_canonical_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(bitmap_dev, lab.as<int>(), rank.as<int>(), HW, n, labels_out_dev);
    if ((rc = check_launch(ctx, "ccl_canonical"))) break;
    if (n_components_host) {
      std::vector<int> r(B + 1);
      for (int b = 0; b <= B && rc == OCRB_OK; ++b) {
        cudaError_t e = cudaMemcpyAsync(&r[b], rank.as<int>() + (int64_t)b * HW, 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) { set_error("memcpy rank: %s", cudaGetErrorString(e)); rc = OCRB_ERR_CUDA; }
      }
      if (rc) break;
      if ((rc = sync(ctx))) break;
      for (int b = 0; b < B; ++b) n_components_host[b] = r[b + 1] - r[b];
    } else {
      rc = sync(ctx);
    }
  } while (0);
  cudaStreamSynchronize(ctx->stream);
  lab.release(); flag.release(); rank.release(); scratch.release();
  return rc;
}

}  // namespace ocrb
