// Polygon -> glyph crop glue ("crop spec v1"): every kept polygon's min-area rectangle is cut out of the source image
// into K cells along its reading axis and each cell is resized to a 28x28 glyph tile for the recognition net.
// The reference has no counterpart (character segmentation is an open item of README.md:20-26; its recognition net is fed
// ready-made 28x28 files, image_ops.rs:73-85), so the definition is ours, built from the reference's own pieces — the
// box of get_min_area_bounding_box (metrics.rs:133-148), the Triangle filter of preprocess_image (image 0.23.11,
// SURVEY A.7) — and stated once in oracle/postproc_oracle.c (orc_crop_glyphs), which this kernel matches bit for bit.
// Compiled with -fmad=false: every f32 operation rounds separately, like the oracle's.
//
// One CTA per glyph.  The vertical pass (rectify by nearest sampling + Triangle filter over the patch rows) writes a
// [28][sw] u8 strip into shared memory, the horizontal pass reduces it to 28x28.  A cell wider than the strip buffer
// is processed one output column at a time over the window of strip columns that column needs.
#include "common.cuh"

namespace ocrb {

constexpr int CROP_THREADS = 256;
constexpr int CROP_STRIP_W = 1024;  // strip columns held in shared memory at a time (28 KB)
constexpr int CROP_FAST_W = 512;    // common case: cell width / patch height bound
constexpr int CROP_MAXT = 16;       // common case: filter taps per output sample

__device__ __forceinline__ float tri_w(float x) {
  const float a = fabsf(x);
  return a < 1.0f ? 1.0f - a : 0.0f;
}

struct CropGeom {
  int ox, oy, pw, ph, W, H;
  float ux, uy, vx, vy;
};

__device__ __forceinline__ uint8_t patch_at(const uint8_t *__restrict__ img, const CropGeom &g, int x, int y) {
  const float s = ((float)x + 0.5f) / (float)g.pw, t = ((float)y + 0.5f) / (float)g.ph;
  const float fx = ((float)g.ox + s * g.ux) + t * g.vx, fy = ((float)g.oy + s * g.uy) + t * g.vy;
  int ix = (int)floorf(fx), iy = (int)floorf(fy);
  ix = ix < 0 ? 0 : (ix > g.W - 1 ? g.W - 1 : ix);
  iy = iy < 0 ? 0 : (iy > g.H - 1 ? g.H - 1 : iy);
  return img[(int64_t)iy * g.W + ix];
}

// filter taps of output sample o when n_in samples become n_out (image 0.23.11 horizontal_sample / vertical_sample)
struct Taps { int left, right; float inputc2, sratio; };
__device__ __forceinline__ Taps taps_for(int o, int n_in, int n_out) {
  const float ratio = (float)n_in / (float)n_out;
  const float sratio = ratio < 1.0f ? 1.0f : ratio;
  const float support = 1.0f * sratio;
  const float inputc = ((float)o + 0.5f) * ratio;
  long long left = (long long)floorf(inputc - support);
  if (left < 0) left = 0;
  if (left > n_in - 1) left = n_in - 1;
  long long right = (long long)ceilf(inputc + support);
  if (right < left + 1) right = left + 1;
  if (right > n_in) right = n_in;
  Taps t;
  t.left = (int)left; t.right = (int)right; t.inputc2 = inputc - 0.5f; t.sratio = sratio;
  return t;
}

__device__ __forceinline__ uint8_t finish_u8(float t, float sum) {
  t = t / sum;
  const float cl = t < 0.0f ? 0.0f : (t > 255.0f ? 255.0f : t);
  return (uint8_t)cl;  // truncation
}

// images: [B][H][W] u8; boxes: [n][4] (TL, TR, BR, BL) in map coordinates; box_image[n]: image of each box (null: image 0)
// out: [n * K][784] u8
__global__ void __launch_bounds__(CROP_THREADS) crop_glyphs_kernel(const uint8_t *__restrict__ images, int H, int W, const int2 *__restrict__ boxes,
                                                                   const int *__restrict__ box_image, int n_boxes, int K, uint8_t *__restrict__ out) {
  __shared__ uint8_t strip[28 * CROP_STRIP_W];
  const int gi = blockIdx.x;
  if (gi >= n_boxes * K) return;
  const int bi = gi / K, cell = gi - bi * K;
  const uint8_t *img = images + (int64_t)(box_image ? box_image[bi] : 0) * H * W;
  const int2 b0 = boxes[bi * 4 + 0], b1 = boxes[bi * 4 + 1], b3 = boxes[bi * 4 + 3];
  CropGeom g;
  g.ox = b0.x; g.oy = b0.y; g.W = W; g.H = H;
  g.ux = (float)(b1.x - b0.x); g.uy = (float)(b1.y - b0.y);
  g.vx = (float)(b3.x - b0.x); g.vy = (float)(b3.y - b0.y);
  float wlen = sqrtf(g.ux * g.ux + g.uy * g.uy), hlen = sqrtf(g.vx * g.vx + g.vy * g.vy);
  if (hlen > wlen) {
    float t;
    t = g.ux; g.ux = g.vx; g.vx = t;
    t = g.uy; g.uy = g.vy; g.vy = t;
    t = wlen; wlen = hlen; hlen = t;
  }
  g.pw = (int)roundf(wlen); g.ph = (int)roundf(hlen);
  if (g.pw < 1) g.pw = 1;
  if (g.ph < 1) g.ph = 1;
  int x0 = (int)((long long)cell * g.pw / K);
  if (x0 > g.pw - 1) x0 = g.pw - 1;
  int x1 = (int)((long long)(cell + 1) * g.pw / K);
  if (x1 > g.pw) x1 = g.pw;
  if (x1 < x0 + 1) x1 = x0 + 1;
  const int sw = x1 - x0;
  uint8_t *tile = out + (int64_t)gi * 784;

  // ---- common case (a cell of at most 14 KB of rectified pixels, filters of at most CROP_MAXT taps): every quantity that
  // depends on one index only is computed once — the two halves of the sampling position per patch column / patch row, the
  // filter weights and their sums per output row / column — and every rectified pixel is sampled once into shared memory.
  // The operations and their order are those of the general path below (and of the oracle), so the results are the same bits.
  if (sw <= CROP_FAST_W && g.ph <= CROP_FAST_W && g.ph * sw <= 14 * 1024) {
    __shared__ float s_ax[CROP_FAST_W], s_ay[CROP_FAST_W], s_bx[CROP_FAST_W], s_by[CROP_FAST_W];
    __shared__ float s_w[2][28][CROP_MAXT], s_sum[2][28];
    __shared__ int s_left[2][28], s_n[2][28];
    uint8_t *patch = strip + 14 * 1024;  // [ph][sw]; the strip itself is [28][sw] here
    int overflow = 0;
    for (int c = threadIdx.x; c < sw; c += CROP_THREADS) {
      const float s = ((float)(x0 + c) + 0.5f) / (float)g.pw;
      s_ax[c] = (float)g.ox + s * g.ux;
      s_ay[c] = (float)g.oy + s * g.uy;
    }
    for (int i = threadIdx.x; i < g.ph; i += CROP_THREADS) {
      const float t = ((float)i + 0.5f) / (float)g.ph;
      s_bx[i] = t * g.vx;
      s_by[i] = t * g.vy;
    }
    if ((threadIdx.x & 31) < 28 && threadIdx.x < 64) {
      const int pass = threadIdx.x >> 5, o = threadIdx.x & 31;  // pass 0: vertical (rows), pass 1: horizontal (columns)
      const Taps tp = taps_for(o, pass ? sw : g.ph, 28);
      const int n = tp.right - tp.left;
      s_left[pass][o] = tp.left;
      s_n[pass][o] = n;
      if (n > CROP_MAXT) {
        overflow = 1;
      } else {
        float sum = 0.0f;
        for (int k = 0; k < n; ++k) {
          const float w = tri_w(((float)(tp.left + k) - tp.inputc2) / tp.sratio);
          sum += w;
          s_w[pass][o][k] = w;
        }
        s_sum[pass][o] = sum;
      }
    }
    if (!__syncthreads_or(overflow)) {
      // four independent image reads in flight per thread (the kernel is bound by their latency)
      const int n_px = g.ph * sw;
      for (int base = threadIdx.x; base < n_px; base += 4 * CROP_THREADS) {
        uint8_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int item = base + u * CROP_THREADS;
          v[u] = 0;
          if (item < n_px) {
            const int i = item / sw, c = item - i * sw;
            const float fx = s_ax[c] + s_bx[i], fy = s_ay[c] + s_by[i];
            int ix = (int)floorf(fx), iy = (int)floorf(fy);
            ix = ix < 0 ? 0 : (ix > g.W - 1 ? g.W - 1 : ix);
            iy = iy < 0 ? 0 : (iy > g.H - 1 ? g.H - 1 : iy);
            v[u] = __ldg(img + (int64_t)iy * g.W + ix);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (base + u * CROP_THREADS < n_px) patch[base + u * CROP_THREADS] = v[u];
      }
      __syncthreads();
      for (int item = threadIdx.x; item < 28 * sw; item += CROP_THREADS) {
        const int r = item / sw, c = item - r * sw;
        const uint8_t *col = patch + s_left[0][r] * sw + c;
        const int n = s_n[0][r];
        float t = 0.0f;
        for (int k = 0; k < n; ++k) t += (float)col[k * sw] * s_w[0][r][k];
        strip[item] = finish_u8(t, s_sum[0][r]);
      }
      __syncthreads();
      for (int item = threadIdx.x; item < 784; item += CROP_THREADS) {
        const int r = item / 28, o = item - r * 28;
        const uint8_t *row = strip + r * sw + s_left[1][o];
        const int n = s_n[1][o];
        float t = 0.0f;
        for (int k = 0; k < n; ++k) t += (float)row[k] * s_w[1][o][k];
        tile[item] = finish_u8(t, s_sum[1][o]);
      }
      return;
    }
  }

  // vertical pass for strip columns [c0, c1) -> strip[r][c - c0]
  auto vertical = [&](int c0, int c1) {
    const int nc = c1 - c0;
    for (int item = threadIdx.x; item < 28 * nc; item += CROP_THREADS) {
      const int r = item / nc, c = item - r * nc;
      const Taps tp = taps_for(r, g.ph, 28);
      float sum = 0.0f, t = 0.0f;
      for (int i = tp.left; i < tp.right; ++i) {
        const float w = tri_w(((float)i - tp.inputc2) / tp.sratio);
        sum += w;
        t += (float)patch_at(img, g, x0 + c0 + c, i) * w;
      }
      strip[r * CROP_STRIP_W + c] = finish_u8(t, sum);
    }
  };
  if (sw <= CROP_STRIP_W) {
    vertical(0, sw);
    __syncthreads();
    for (int item = threadIdx.x; item < 784; item += CROP_THREADS) {
      const int r = item / 28, o = item - r * 28;
      const Taps tp = taps_for(o, sw, 28);
      float sum = 0.0f, t = 0.0f;
      for (int i = tp.left; i < tp.right; ++i) {
        const float w = tri_w(((float)i - tp.inputc2) / tp.sratio);
        sum += w;
        t += (float)strip[r * CROP_STRIP_W + i] * w;
      }
      tile[item] = finish_u8(t, sum);
    }
  } else {
    // very wide cell: one output column at a time over the window of strip columns it needs (the window is
    // 2 * sw / 28 + 2 columns wide; a window beyond the buffer — a cell wider than ~14000 px — is clipped to it)
    for (int o = 0; o < 28; ++o) {
      const Taps tp = taps_for(o, sw, 28);
      const int c1 = tp.right - tp.left > CROP_STRIP_W ? tp.left + CROP_STRIP_W : tp.right;
      __syncthreads();
      vertical(tp.left, c1);
      __syncthreads();
      for (int r = threadIdx.x; r < 28; r += CROP_THREADS) {
        float sum = 0.0f, t = 0.0f;
        for (int i = tp.left; i < c1; ++i) {
          const float w = tri_w(((float)i - tp.inputc2) / tp.sratio);
          sum += w;
          t += (float)strip[r * CROP_STRIP_W + (i - tp.left)] * w;
        }
        tile[r * 28 + o] = finish_u8(t, sum);
      }
    }
  }
}

int launch_crop_glyphs(ocrb_ctx *ctx, const uint8_t *images, int H, int W, const int2 *boxes, const int *box_image, int n_boxes, int K,
                       uint8_t *out) {
  if (n_boxes <= 0 || K <= 0) return OCRB_OK;
  crop_glyphs_kernel<<<(unsigned)((int64_t)n_boxes * K), CROP_THREADS, 0, ctx->stream>>>(images, H, W, boxes, box_image, n_boxes, K, out);
  return check_launch(ctx, "crop_glyphs");
}

}  // namespace ocrb

using namespace ocrb;

// test hook of the crop stage: one image, n boxes (TL, TR, BR, BL as ocrb_min_area_bounding_box returns them)
extern "C" int ocrb_crop_glyphs(ocrb_ctx *ctx, const uint8_t *image, int H, int W, const int32_t *boxes_xy, int n_boxes, int glyphs_per_box,
                                uint8_t *out_glyphs) {
  OCRB_REQUIRE(ctx && image && boxes_xy && out_glyphs && H > 0 && W > 0 && n_boxes > 0 && glyphs_per_box > 0, "bad argument");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  const void *img_dev = nullptr, *box_dev = nullptr;
  void *out_dev = nullptr;
  const size_t out_bytes = (size_t)n_boxes * glyphs_per_box * 784;
  OCRB_TRY(to_device(ctx, 0, image, (size_t)H * W, &img_dev));
  OCRB_TRY(to_device(ctx, 1, boxes_xy, (size_t)n_boxes * 32, &box_dev));
  OCRB_TRY(out_device(ctx, 2, out_glyphs, out_bytes, &out_dev));
  OCRB_TRY(launch_crop_glyphs(ctx, (const uint8_t *)img_dev, H, W, (const int2 *)box_dev, nullptr, n_boxes, glyphs_per_box, (uint8_t *)out_dev));
  OCRB_TRY(finish_output(ctx, out_glyphs, out_dev, out_bytes));
  return sync(ctx);
}
