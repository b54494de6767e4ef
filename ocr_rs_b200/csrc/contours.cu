// Border following on the GPU: where every Suzuki–Abe border of imageproc::find_contours
// (metrics.rs:78-81) starts, the chains themselves, and their Douglas–Peucker polygons
// (imageproc::approximate_polygon_dp at metrics.rs:87-95).
//
// The sequential sign-marking scan is replaced by facts derived from the component labels
// (ccl.cu) — verified against the sequential algorithm on random images (tests/):
//   * a foreground component whose raster-first pixel p0 has x > 0 gets ONE outer border,
//     started at p0 coming from the west;
//   * every background component that does not reach the image frame gets ONE hole border,
//     started at the pixel west of its raster-first pixel, coming from the east;
//   * a foreground component whose raster-first pixel lies in column 0 ("left-anchored":
//     imageproc's `x > 0` guard suppresses the outer start there) is replayed sequentially,
//     crack by crack, by one warp: the marks of the original algorithm reduce to "which
//     borders of this component have been traced so far";
//   * a chain is a pure function of (bitmap, start pixel, start direction).
// Contours are emitted in raster order of their start pixel, like the reference.
#include "ccl.cuh"
#include <vector>

#include "common.cuh"
#include "scan.cuh"

namespace ocrb {

enum : uint8_t { START_NONE = 0, START_OUTER = 1, START_HOLE = 2 };

// ---------------------------------------------------------------------------------------
// The labels arrive as an un-flattened union-find forest whose roots are final.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int ccl_find(const int *__restrict__ L, int a) {
  int p = L[a];
  while (p != a) {
    a = p;
    p = L[a];
  }
  return a;
}

// frame pixels only: background components reaching the frame are "open" (no hole border);
// a foreground root in column 0 means a left-anchored component exists (slow path needed)
__global__ void contour_frame_kernel(const uint8_t *__restrict__ bitmap, const int *__restrict__ labels, int H, int W, int B,
                                     uint8_t *__restrict__ bg_open, int *__restrict__ need_anchored, CclTiles tiles) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int per = 2 * W + 2 * H;
  if (t >= (int64_t)per * B) return;
  const int b = (int)(t / per), k = (int)(t % per);
  int x, y;
  if (k < W) { x = k; y = 0; }
  else if (k < 2 * W) { x = k - W; y = H - 1; }
  else if (k < 2 * W + H) { x = 0; y = k - 2 * W; }
  else { x = W - 1; y = k - 2 * W - H; }
  const int64_t HW = (int64_t)H * W;
  const int i = y * W + x;
  const int *L = labels + b * HW;
  if (bitmap[b * HW + i] == 0) bg_open[b * HW + ccl_find_px(L, tiles, b, i, W)] = 1;  // a background pixel may sit in a tile that stores no labels
  else if (x == 0 && L[i] == i) *need_anchored = 1;
}

// bounding boxes of left-anchored foreground components (keyed by the row of their root, which
// is in column 0); does nothing unless such a component exists
__global__ void contour_props_kernel(const uint8_t *__restrict__ bitmap, const int *__restrict__ labels, int H, int W,
                                     int B, const int *__restrict__ need_anchored, int4 *__restrict__ anchored_bbox) {
  if (*need_anchored == 0) return;
  const int64_t HW = (int64_t)H * W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < HW * B; idx += (int64_t)gridDim.x * blockDim.x) {
    if (bitmap[idx] == 0) continue;
    const int64_t b = idx / HW;
    const int i = (int)(idx % HW);
    const int root = ccl_find(labels + b * HW, i);
    if (root % W == 0) {
      int4 *bb = anchored_bbox + b * H + root / W;
      atomicMin(&bb->x, i % W);
      atomicMax(&bb->y, i % W);
      atomicMin(&bb->z, i / W);
      atomicMax(&bb->w, i / W);
    }
  }
}

__global__ void contour_bbox_init_kernel(int4 *bbox, int64_t n, int *need_anchored) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) bbox[i] = make_int4(INT32_MAX, -1, INT32_MAX, -1);
  if (i == 0) *need_anchored = 0;
}

// closed-form starts for ordinary components: outer = a foreground root outside column 0;
// hole = the pixel west of a closed background component's root
__global__ void contour_start_flags_kernel(const uint8_t *__restrict__ bitmap, const int *__restrict__ labels, int H, int W,
                                           int B, const uint8_t *__restrict__ bg_open, uint8_t *__restrict__ flags, CclTiles tiles) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t HW = (int64_t)H * W;
  if (idx >= HW * B) return;
  int i = (int)(idx % HW);
  int x = i % W;
  uint8_t f = START_NONE;
  if (bitmap[idx] != 0) {
    if (labels[idx] == i) {
      if (x != 0) f = START_OUTER;  // a root in column 0 is left-anchored: replayed by contour_anchored_kernel
    } else if (x + 1 < W && bitmap[idx + 1] == 0 && ccl_parent(labels + (idx - i), tiles, (int)(idx / HW), i + 1, W) == i + 1 && !bg_open[idx + 1]) {
      const int root = ccl_find(labels + (idx - i), i);
      if (root % W != 0) f = START_HOLE;
    }
  }
  flags[idx] = f;
}

// same, 4 pixels per thread (W % 4 == 0): one 32-bit bitmap load decides whether the labels are
// needed at all, and the flags go out as one 32-bit store
__global__ void contour_start_flags_vec4_kernel(const uint8_t *__restrict__ bitmap, const int *__restrict__ labels, int H, int W,
                                                int B, const uint8_t *__restrict__ bg_open, uint8_t *__restrict__ flags, CclTiles tiles) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t HW = (int64_t)H * W;
  if (q * 4 >= HW * B) return;
  const int64_t idx0 = q * 4;
  const uint32_t bm = *reinterpret_cast<const uint32_t *>(bitmap + idx0);
  uint32_t out = 0;
  if (bm != 0) {
    const int i0 = (int)(idx0 % HW);
    const int x0 = i0 % W;
    const int4 lab = *reinterpret_cast<const int4 *>(labels + idx0);
    const int labs[4] = {lab.x, lab.y, lab.z, lab.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (((bm >> (8 * k)) & 0xffu) == 0) continue;
      const int i = i0 + k, x = x0 + k;
      uint32_t f = START_NONE;
      if (labs[k] == i) {
        if (x != 0) f = START_OUTER;
      } else if (x + 1 < W) {
        const bool east_bg = k < 3 ? ((bm >> (8 * (k + 1))) & 0xffu) == 0 : bitmap[idx0 + 4] == 0;
        if (east_bg) {
          // inside this 4-pixel group the east pixel shares the tile of a foreground pixel; across groups it may not
          const int east_lab = k < 3 ? labs[k + 1] : ccl_parent(labels + (idx0 - i0), tiles, (int)(idx0 / HW), i + 1, W);
          if (east_lab == i + 1 && !bg_open[idx0 + k + 1]) {
            const int root = ccl_find(labels + (idx0 - i0), i);
            if (root % W != 0) f = START_HOLE;
          }
        }
      }
      out |= f << (8 * k);
    }
  }
  *reinterpret_cast<uint32_t *>(flags + idx0) = out;
}

// same, 16 pixels per thread (W % 16 == 0) and one scan tile (SCAN_TILE pixels) per block: most 16-pixel groups of a
// page are empty (one 16-byte load, one 16-byte store), and the block's count of starts goes straight into the
// per-tile counts of the order-preserving compaction — the flags are not re-read to be counted.
constexpr int FLAGS16_THREADS = SCAN_TILE / 16;
__global__ void __launch_bounds__(FLAGS16_THREADS) contour_start_flags_vec16_kernel(const uint8_t *__restrict__ bitmap, const int *__restrict__ labels,
                                                                                    int H, int W, int B, const uint8_t *__restrict__ bg_open,
                                                                                    uint8_t *__restrict__ flags, int *__restrict__ tile_counts, CclTiles tiles) {
  __shared__ int s_cnt[FLAGS16_THREADS / 32];
  const int64_t idx0 = ((int64_t)blockIdx.x * FLAGS16_THREADS + threadIdx.x) * 16;
  const int64_t HW = (int64_t)H * W;
  int cnt = 0;
  if (idx0 < HW * B) {
    const uint4 bm4 = *reinterpret_cast<const uint4 *>(bitmap + idx0);
    const uint32_t bmw[4] = {bm4.x, bm4.y, bm4.z, bm4.w};
    uint32_t outw[4] = {0u, 0u, 0u, 0u};
    if ((bm4.x | bm4.y | bm4.z | bm4.w) != 0) {
      const int i0 = (int)(idx0 % HW);
      const int x0 = i0 % W;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t bm = bmw[w];
        if (bm == 0) continue;
        const int4 lab = *reinterpret_cast<const int4 *>(labels + idx0 + 4 * w);
        const int labs[4] = {lab.x, lab.y, lab.z, lab.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (((bm >> (8 * k)) & 0xffu) == 0) continue;
          const int i = i0 + 4 * w + k, x = x0 + 4 * w + k;
          uint32_t f = START_NONE;
          if (labs[k] == i) {
            if (x != 0) f = START_OUTER;
          } else if (x + 1 < W) {
            const int64_t e = idx0 + 4 * w + k + 1;  // east neighbour
            const bool east_bg = k < 3 ? ((bm >> (8 * (k + 1))) & 0xffu) == 0 : (w < 3 ? (bmw[(w + 1) & 3] & 0xffu) == 0 : bitmap[e] == 0);
            if (east_bg) {
              // a 4-pixel group never straddles a tile, so labs[] are stored labels; the next group may lie in a tile that stores none
              const int east_lab = k < 3 ? labs[k + 1] : ccl_parent(labels + (idx0 - i0), tiles, (int)(idx0 / HW), i + 1, W);
              if (east_lab == i + 1 && !bg_open[e]) {
                const int root = ccl_find(labels + (idx0 - i0), i);
                if (root % W != 0) f = START_HOLE;
              }
            }
          }
          outw[w] |= f << (8 * k);
        }
      }
    }
    *reinterpret_cast<uint4 *>(flags + idx0) = make_uint4(outw[0], outw[1], outw[2], outw[3]);
#pragma unroll
    for (int w = 0; w < 4; ++w) cnt += __popc((outw[w] | (outw[w] >> 1)) & 0x01010101u);  // flag values are 0, 1, 2
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int k = 0; k < FLAGS16_THREADS / 32; ++k) t += s_cnt[k];
    tile_counts[blockIdx.x] = t;
  }
}

// sequential replay for left-anchored components; one warp per image row that holds a root
// in column 0.  `hole_traced` is a zero-initialised byte per pixel (indexed by bg root).
__global__ void __launch_bounds__(128) contour_anchored_kernel(const uint8_t *__restrict__ bitmap, const int *__restrict__ labels,
                                                               int H, int W, int B, const uint8_t *__restrict__ bg_open,
                                                               const int4 *__restrict__ anchored_bbox, const int *__restrict__ need_anchored,
                                                               uint8_t *__restrict__ hole_traced, uint8_t *__restrict__ flags,
                                                               int *__restrict__ tile_counts /* per-SCAN_TILE start counts to keep current, or null */,
                                                               CclTiles tiles) {
  if (*need_anchored == 0) return;
  const int lane = threadIdx.x & 31;
  int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= (int64_t)B * H) return;
  int64_t b = wid / H;
  int row = (int)(wid % H);
  int64_t HW = (int64_t)H * W;
  const uint8_t *bm = bitmap + b * HW;
  const int *L = labels + b * HW;
  const uint8_t *open = bg_open + b * HW;
  volatile uint8_t *traced = hole_traced + b * HW;
  uint8_t *fl = flags + b * HW;
  const int F = row * W;
  if (bm[F] == 0 || L[F] != F) return;  // warp-uniform
  int4 bb = anchored_bbox[b * H + row];
  bool traced_inf = false;
  for (int y = bb.z; y <= bb.w; ++y) {
    for (int xb = bb.x; xb <= bb.y; xb += 32) {
      int x = xb + lane;
      bool in_f = x <= bb.y && bm[y * W + x] != 0 && ccl_find(L, y * W + x) == F;
      bool wcr = in_f && x > 0 && bm[y * W + x - 1] == 0;
      bool ecr = in_f && x + 1 < W && bm[y * W + x + 1] == 0;
      uint32_t cand = __ballot_sync(0xffffffffu, wcr || ecr);
      uint32_t wmask = __ballot_sync(0xffffffffu, wcr);
      uint32_t emask = __ballot_sync(0xffffffffu, ecr);
      while (cand) {
        int l = __ffs(cand) - 1;
        cand &= cand - 1;
        int qx = xb + l, q = y * W + qx;
        // labels of the 4-neighbour background pixels (-1 = outer background / out of image)
        int lab[4];
        const int dx[4] = {-1, 1, 0, 0}, dy[4] = {0, 0, -1, 1};
        bool visited = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          int nx = qx + dx[k], ny = y + dy[k];
          lab[k] = -2;  // not background
          if (nx < 0 || ny < 0 || nx >= W || ny >= H) lab[k] = -1;
          else if (bm[ny * W + nx] == 0) {
            int r = ccl_find_px(L, tiles, (int)b, ny * W + nx, W);
            lab[k] = open[r] ? -1 : r;
          }
          if (lab[k] == -1) visited |= traced_inf;
          else if (lab[k] >= 0) visited |= (traced[lab[k]] != 0);
        }
        bool has_w = (wmask >> l) & 1, has_e = (emask >> l) & 1;
        if (has_w && !visited) {
          if (lab[0] == -1) traced_inf = true;
          else if (lane == 0) traced[lab[0]] = 1;
          if (lane == 0) {
            fl[q] = START_OUTER;  // anchored starts only ever land on pixels the closed-form pass left at NONE
            if (tile_counts) atomicAdd(&tile_counts[(b * HW + q) / SCAN_TILE], 1);
          }
        } else if (has_e) {
          bool e_traced = lab[1] == -1 ? traced_inf : (traced[lab[1]] != 0);
          if (!e_traced) {
            if (lab[1] == -1) traced_inf = true;
            else if (lane == 0) traced[lab[1]] = 1;
            if (lane == 0) {
              fl[q] = START_HOLE;
              if (tile_counts) atomicAdd(&tile_counts[(b * HW + q) / SCAN_TILE], 1);
            }
          }
        }
        __syncwarp();
      }
    }
  }
}

// flags -> contour records (start pixel as batch-global index, kind), in raster order:
// order-preserving compaction with per-tile counts (scan.cuh tile reduce + scan of the tile
// sums) and an in-tile rank computed here — no per-pixel offset array is ever written.
__global__ void __launch_bounds__(SCAN_THREADS) contour_records_kernel(const uint8_t *__restrict__ flags, const int *__restrict__ tile_offs,
                                                                       int64_t n, int64_t *__restrict__ start_idx, uint8_t *__restrict__ kind) {
  __shared__ int smem[33];
  // exclusive tile offsets with the grand total in the slot after the last tile: an empty tile (most of a page) is skipped unread
  if (tile_offs[blockIdx.x + 1] == tile_offs[blockIdx.x]) return;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint8_t f[SCAN_ITEMS];
  int s = 0;
  if (base + SCAN_ITEMS <= n) {
    const uint2 u = *reinterpret_cast<const uint2 *>(flags + base);  // SCAN_ITEMS == 8, base % 8 == 0
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[j] = (u.x >> (8 * j)) & 0xff; f[4 + j] = (u.y >> (8 * j)) & 0xff; }
  } else {
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) f[j] = base + j < n ? flags[base + j] : 0;
  }
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) s += f[j] != 0;
  int total;
  int rank = block_exclusive_scan<int>(s, &total, smem) + tile_offs[blockIdx.x];
  if (s == 0) return;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j)
    if (f[j]) {
      start_idx[rank] = base + j;
      kind[rank] = f[j] - 1;
      ++rank;
    }
}

// ---------------------------------------------------------------------------------------
// chain tracing.  Neighbour ring, clockwise on screen: W NW N NE E SE S SW.
// ---------------------------------------------------------------------------------------
// nibble tables: dx+1 for d=0..7 = {0,0,1,2,2,2,1,0}; dy+1 = {1,0,0,0,1,2,2,2}
__device__ __forceinline__ void ring(int d, int &dx, int &dy) {
  dx = (int)((0x01222100u >> (d * 4)) & 0xf) - 1;
  dy = (int)((0x22210001u >> (d * 4)) & 0xf) - 1;
}

struct Tracer {
  const uint8_t *bm;
  int W, H;
  __device__ __forceinline__ bool nz(int x, int y) const {
    return x >= 0 && y >= 0 && x < W && y < H && bm[y * W + x] != 0;
  }
};

// Walks one border; calls emit(x, y) for every chain point; returns the chain length.
template <class Emit>
__device__ __forceinline__ int64_t trace_border(const Tracer &t, int sx, int sy, int from, Emit emit) {
  int d1 = -1;
#pragma unroll 1
  for (int k = 0; k < 8; ++k) {
    int d = (from + k) & 7, dx, dy;
    ring(d, dx, dy);
    if (t.nz(sx + dx, sy + dy)) { d1 = d; break; }
  }
  if (d1 < 0) {
    emit(sx, sy);
    return 1;
  }
  int dx, dy;
  ring(d1, dx, dy);
  const int p1x = sx + dx, p1y = sy + dy;
  int p3x = sx, p3y = sy;
  int dp2 = d1;  // direction from p3 to p2
  int64_t n = 0;
  for (;;) {
    emit(p3x, p3y);
    ++n;
    int d4 = dp2;
#pragma unroll 1
    for (int k = 1; k <= 8; ++k) {
      int d = (dp2 - k) & 7;
      ring(d, dx, dy);
      if (t.nz(p3x + dx, p3y + dy)) { d4 = d; break; }
    }
    ring(d4, dx, dy);
    int p4x = p3x + dx, p4y = p3y + dy;
    if (p4x == sx && p4y == sy && p3x == p1x && p3y == p1y) break;
    // next step: p2 <- p3, p3 <- p4; direction from new p3 back to new p2 is the opposite of d4
    dp2 = (d4 + 4) & 7;
    p3x = p4x;
    p3y = p4y;
  }
  return n;
}

__global__ void trace_count_kernel(const uint8_t *__restrict__ bitmap, int H, int W, const int64_t *__restrict__ start_idx,
                                   const uint8_t *__restrict__ kind, int64_t n_contours, int *__restrict__ lengths) {
  const int64_t c = sparse_item_index();
  if (c < 0 || c >= n_contours) return;
  int64_t HW = (int64_t)H * W;
  int64_t g = start_idx[c];
  int64_t b = g / HW;
  int i = (int)(g % HW);
  Tracer t{bitmap + b * HW, W, H};
  int64_t n = trace_border(t, i % W, i / W, kind[c] ? 4 : 0, [](int, int) {});
  lengths[c] = (int)n;
}

__global__ void trace_store_kernel(const uint8_t *__restrict__ bitmap, int H, int W, const int64_t *__restrict__ start_idx,
                                   const uint8_t *__restrict__ kind, int64_t n_contours,
                                   const int64_t *__restrict__ chain_off, ushort2 *__restrict__ chain) {
  const int64_t c = sparse_item_index();
  if (c < 0 || c >= n_contours) return;
  int64_t HW = (int64_t)H * W;
  int64_t g = start_idx[c];
  int64_t b = g / HW;
  int i = (int)(g % HW);
  Tracer t{bitmap + b * HW, W, H};
  ushort2 *out = chain + chain_off[c];
  int64_t k = 0;
  trace_border(t, i % W, i / W, kind[c] ? 4 : 0, [&](int x, int y) { out[k++] = make_ushort2((unsigned short)x, (unsigned short)y); });
}

// ---------------------------------------------------------------------------------------
// Douglas–Peucker exactly as imageproc does it, iteratively: the kept points
// are the left ends of the leaf ranges plus the chain end; closed => the last one is
// popped; metrics.rs:92-94 pops once more if first == last.  f64 arithmetic, first index
// of the strictly largest distance, NaN (coincident range ends) never splits.
// One thread per contour; `stack` shares the chain's arena offsets.
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ double dist_d(double ax, double ay, double bx, double by) {
  double dx = ax - bx, dy = ay - by;
  return sqrt(dx * dx + dy * dy);
}

__global__ void approx_dp_kernel(const ushort2 *__restrict__ chain, const int64_t *__restrict__ chain_off,
                                 int64_t n_contours, int *__restrict__ stack, ushort2 *__restrict__ dp_out,
                                 int *__restrict__ dp_count) {
  const int64_t c = sparse_item_index();
  if (c < 0 || c >= n_contours) return;
  const int64_t off = chain_off[c];
  const int n = (int)(chain_off[c + 1] - off);
  const ushort2 *p = chain + off;
  ushort2 *out = dp_out + off;
  // [approx-dp-body-begin] (repeated verbatim in approx_polygon_host below; tests/test_clip_polygon_host.py keeps the two in step)
  if (n == 1) {  // the open chain {p0, p0} minus the closing pop (the general loop would write 2 slots)
    out[0] = p[0];
    dp_count[c] = 1;
    return;
  }
  // arc_length(closed = true)
  double len = 0.0;
  for (int i = 0; i + 1 < n; ++i) len += dist_d(p[i].x, p[i].y, p[i + 1].x, p[i + 1].y);
  if (n > 2) len += dist_d(p[0].x, p[0].y, p[n - 1].x, p[n - 1].y);
  double eps = 0.01 * len;
  if (eps == 0.) eps = 0.01;
  int *st = stack + off;  // holds the pending right ends
  int sp = 0, m = 0;
  int lo = 0, hi = n - 1;
  for (;;) {
    double x0 = p[lo].x, y0 = p[lo].y, x1 = p[hi].x, y1 = p[hi].y;
    double a = y0 - y1, bq = x1 - x0, cc = x0 * y1 - x1 * y0;
    double den = sqrt(a * a + bq * bq);
    double dmax = 0.0;
    int index = lo;
    for (int i = lo + 1; i <= hi; ++i) {
      double d = fabs(a * (double)p[i].x + bq * (double)p[i].y + cc) / den;
      if (d > dmax) { index = i; dmax = d; }
    }
    if (dmax > eps) {
      st[sp++] = hi;  // right part [index, hi] waits
      hi = index;
      continue;
    }
    out[m++] = p[lo];  // leaf range [lo, hi]
    if (sp == 0) {
      out[m++] = p[hi];
      break;
    }
    lo = hi;
    hi = st[--sp];
  }
  m -= 1;  // closed => pop
  if (m > 1 && out[0].x == out[m - 1].x && out[0].y == out[m - 1].y) m -= 1;
  dp_count[c] = m;
  // [approx-dp-body-end]
}

// host-only test hook (ocrb_debug_approx_polygon_host): approximate_polygon_dp as metrics.rs:87-95 uses it, run on the HOST
// by the kernel's own statements — the block between the markers is the kernel's, verbatim (a CPU test compares the two
// texts), so the CPU suite holds the shipped algorithm to the oracle without a device and without touching the kernel.
int approx_polygon_host(const int32_t *chain_xy, int64_t n_pts, int32_t *out_xy, int64_t out_cap_pts, int64_t *n_out) {
  std::vector<ushort2> pv((size_t)n_pts), outv((size_t)n_pts + 1);
  std::vector<int> stackv((size_t)n_pts + 1);
  for (int64_t i = 0; i < n_pts; ++i) pv[i] = make_ushort2((unsigned short)chain_xy[2 * i], (unsigned short)chain_xy[2 * i + 1]);
  int count[1] = {0};
  auto run = [](const ushort2 *p, int n, ushort2 *out, int *stack, int64_t off, int *dp_count, int c) {
  // [approx-dp-body-begin]
  if (n == 1) {  // the open chain {p0, p0} minus the closing pop (the general loop would write 2 slots)
    out[0] = p[0];
    dp_count[c] = 1;
    return;
  }
  // arc_length(closed = true)
  double len = 0.0;
  for (int i = 0; i + 1 < n; ++i) len += dist_d(p[i].x, p[i].y, p[i + 1].x, p[i + 1].y);
  if (n > 2) len += dist_d(p[0].x, p[0].y, p[n - 1].x, p[n - 1].y);
  double eps = 0.01 * len;
  if (eps == 0.) eps = 0.01;
  int *st = stack + off;  // holds the pending right ends
  int sp = 0, m = 0;
  int lo = 0, hi = n - 1;
  for (;;) {
    double x0 = p[lo].x, y0 = p[lo].y, x1 = p[hi].x, y1 = p[hi].y;
    double a = y0 - y1, bq = x1 - x0, cc = x0 * y1 - x1 * y0;
    double den = sqrt(a * a + bq * bq);
    double dmax = 0.0;
    int index = lo;
    for (int i = lo + 1; i <= hi; ++i) {
      double d = fabs(a * (double)p[i].x + bq * (double)p[i].y + cc) / den;
      if (d > dmax) { index = i; dmax = d; }
    }
    if (dmax > eps) {
      st[sp++] = hi;  // right part [index, hi] waits
      hi = index;
      continue;
    }
    out[m++] = p[lo];  // leaf range [lo, hi]
    if (sp == 0) {
      out[m++] = p[hi];
      break;
    }
    lo = hi;
    hi = st[--sp];
  }
  m -= 1;  // closed => pop
  if (m > 1 && out[0].x == out[m - 1].x && out[0].y == out[m - 1].y) m -= 1;
  dp_count[c] = m;
  // [approx-dp-body-end]
  };
  run(pv.data(), (int)n_pts, outv.data(), stackv.data(), 0, count, 0);
  *n_out = count[0];
  if (count[0] > out_cap_pts) { set_error("approx_polygon: %d points, room for %lld", count[0], (long long)out_cap_pts); return OCRB_ERR_CAPACITY; }
  for (int i = 0; i < count[0]; ++i) { out_xy[2 * i] = outv[i].x; out_xy[2 * i + 1] = outv[i].y; }
  return OCRB_OK;
}

// ---------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------
int launch_contour_starts(ocrb_ctx *ctx, const uint8_t *bitmap, const int *labels, int B, int H, int W,
                          uint8_t *bg_open /*B*HW, zeroed here*/, uint8_t *hole_traced /*B*HW, zeroed here*/,
                          int4 *anchored_bbox /*B*H*/, uint8_t *flags /*B*HW*/, int *need_anchored /*device int*/,
                          int *tile_counts /* cdiv(B*HW, SCAN_TILE) ints: start flags per scan tile */) {
  int64_t n = (int64_t)B * H * W;
  // the labelling that produced `labels` left its per-tile "no foreground" flags in the ctx (launch_ccl)
  const int tiles_x = (int)cdiv(W, CCL_TW), tiles_y = (int)cdiv(H, CCL_TH);
  const CclTiles tiles = {ctx->ccl_tile_empty.as<uint8_t>(), tiles_x, tiles_x * tiles_y};
  OCRB_CUDA(cudaMemsetAsync(bg_open, 0, n, ctx->stream));
  OCRB_CUDA(cudaMemsetAsync(hole_traced, 0, n, ctx->stream));
  contour_bbox_init_kernel<<<(unsigned)cdiv((int64_t)B * H, 256), 256, 0, ctx->stream>>>(anchored_bbox, (int64_t)B * H, need_anchored);
  OCRB_TRY(check_launch(ctx, "contour_bbox_init"));
  contour_frame_kernel<<<(unsigned)cdiv((int64_t)B * (2 * W + 2 * H), 256), 256, 0, ctx->stream>>>(bitmap, labels, H, W, B, bg_open, need_anchored, tiles);
  OCRB_TRY(check_launch(ctx, "contour_frame"));
  contour_props_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(bitmap, labels, H, W, B, need_anchored, anchored_bbox);
  OCRB_TRY(check_launch(ctx, "contour_props"));
  const bool counted = W % 16 == 0;  // the flags kernel counts per tile itself
  if (counted)
    contour_start_flags_vec16_kernel<<<(unsigned)cdiv(n, SCAN_TILE), FLAGS16_THREADS, 0, ctx->stream>>>(bitmap, labels, H, W, B, bg_open, flags,
                                                                                                       tile_counts, tiles);
  else if (W % 4 == 0)
    contour_start_flags_vec4_kernel<<<(unsigned)cdiv(n / 4, 256), 256, 0, ctx->stream>>>(bitmap, labels, H, W, B, bg_open, flags, tiles);
  else
    contour_start_flags_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(bitmap, labels, H, W, B, bg_open, flags, tiles);
  OCRB_TRY(check_launch(ctx, "contour_start_flags"));
  int64_t warps = (int64_t)B * H;
  contour_anchored_kernel<<<(unsigned)cdiv(warps * 32, 128), 128, 0, ctx->stream>>>(bitmap, labels, H, W, B, bg_open,
                                                                                    anchored_bbox, need_anchored, hole_traced, flags,
                                                                                    counted ? tile_counts : nullptr, tiles);
  OCRB_TRY(check_launch(ctx, "contour_anchored"));
  if (!counted) {
    scan_tile_reduce_kernel<uint8_t, int, ScanNonZero><<<(unsigned)cdiv(n, SCAN_TILE), SCAN_THREADS, 0, ctx->stream>>>(flags, n, tile_counts);
    OCRB_TRY(check_launch(ctx, "scan_tile_reduce"));
  }
  return OCRB_OK;
}

// phase 1: exclusive scan of the per-tile start counts left by launch_contour_starts; tile_offs must hold
// scan_scratch_elems(n) * 2 ints.  The total lands at tile_offs[tiles] (see contour_count_slot).
int launch_contour_count(ocrb_ctx *ctx, int64_t n, int *tile_offs) {
  const int64_t tiles = cdiv(n, SCAN_TILE);
  int *lvl2 = tile_offs + tiles + 8;
  return exclusive_scan<int, int, ScanIdentity>(ctx, tile_offs, tiles, tile_offs, lvl2);
}
int64_t contour_count_slot(int64_t n) { return cdiv(n, SCAN_TILE); }

int launch_contour_records(ocrb_ctx *ctx, const uint8_t *flags, const int *tile_offs, int64_t n, int64_t *start_idx, uint8_t *kind) {
  contour_records_kernel<<<(unsigned)cdiv(n, SCAN_TILE), SCAN_THREADS, 0, ctx->stream>>>(flags, tile_offs, n, start_idx, kind);
  return check_launch(ctx, "contour_records");
}

int launch_trace_count(ocrb_ctx *ctx, const uint8_t *bitmap, int H, int W, const int64_t *start_idx, const uint8_t *kind,
                       int64_t n_contours, int *lengths) {
  if (n_contours <= 0) return OCRB_OK;
  trace_count_kernel<<<sparse_grid(n_contours, 128), 128, 0, ctx->stream>>>(bitmap, H, W, start_idx, kind, n_contours, lengths);
  return check_launch(ctx, "trace_count");
}

int launch_trace_store(ocrb_ctx *ctx, const uint8_t *bitmap, int H, int W, const int64_t *start_idx, const uint8_t *kind,
                       int64_t n_contours, const int64_t *chain_off, ushort2 *chain) {
  if (n_contours <= 0) return OCRB_OK;
  trace_store_kernel<<<sparse_grid(n_contours, 128), 128, 0, ctx->stream>>>(bitmap, H, W, start_idx, kind, n_contours, chain_off, chain);
  return check_launch(ctx, "trace_store");
}

int launch_approx_dp(ocrb_ctx *ctx, const ushort2 *chain, const int64_t *chain_off, int64_t n_contours, int *stack,
                     ushort2 *dp_out, int *dp_count) {
  if (n_contours <= 0) return OCRB_OK;
  approx_dp_kernel<<<sparse_grid(n_contours, 128), 128, 0, ctx->stream>>>(chain, chain_off, n_contours, stack, dp_out, dp_count);
  return check_launch(ctx, "approx_dp");
}

}  // namespace ocrb
