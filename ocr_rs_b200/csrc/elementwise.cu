// HBM-bound element-wise kernels of the path: binarize, u8<->f32 conversions, and the
// preprocess_image resize/luma/pad.  Compiled with -fmad=false: the resize arithmetic must
// round exactly like the reference's scalar f32 code (image 0.23.11).
//
// reference: metrics.rs:129-131 (binarize), image_ops.rs:350-381 (conversions),
//            image_ops.rs:73-85 (/255), image_ops.rs:188-220 (preprocess_image)
#include <cuda_bf16.h>

#include "common.cuh"

namespace ocrb {

// ---------------------------------------------------------------------------------------
// binarize: out = pred > (float)thresh.   5 B / pixel (4 read + 1 write).
// Vector body: each lane reads float4 (coalesced 512 B per warp request) and writes one
// packed uchar4; 4 independent requests in flight per lane.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) binarize_vec_kernel(const float4 *__restrict__ in, uint32_t *__restrict__ out,
                                                           int64_t n4, float t) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t base = ((int64_t)blockIdx.x * blockDim.x) * 4 + threadIdx.x; base < n4; base += stride) {
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t i = base + (int64_t)j * blockDim.x;
      if (i < n4) v[j] = __ldcs(in + i);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t i = base + (int64_t)j * blockDim.x;
      if (i < n4) {
        uint32_t r = (v[j].x > t ? 1u : 0u) | (v[j].y > t ? 0x100u : 0u) | (v[j].z > t ? 0x10000u : 0u) |
                     (v[j].w > t ? 0x1000000u : 0u);
        out[i] = r;
      }
    }
  }
}

__global__ void binarize_scalar_kernel(const float *__restrict__ in, uint8_t *__restrict__ out, int64_t n, float t) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = in[i] > t ? 1 : 0;
}

int launch_binarize(ocrb_ctx *ctx, const float *pred, int64_t n, float t, uint8_t *out) {
  if (n <= 0) return OCRB_OK;
  bool aligned = ((uintptr_t)pred % 16 == 0) && ((uintptr_t)out % 4 == 0);
  int64_t n4 = aligned ? n / 4 : 0;
  if (n4 > 0) {
    int64_t blocks = cdiv(n4, 256 * 4);
    int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    binarize_vec_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>((const float4 *)pred, (uint32_t *)out, n4, t);
    OCRB_TRY(check_launch(ctx, "binarize_vec"));
  }
  int64_t rem = n - n4 * 4;
  if (rem > 0) {
    int64_t blocks = cdiv(rem, 256);
    if (blocks > 1184) blocks = 1184;
    binarize_scalar_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(pred + n4 * 4, out + n4 * 4, rem, t);
    OCRB_TRY(check_launch(ctx, "binarize_scalar"));
  }
  return OCRB_OK;
}

// ---------------------------------------------------------------------------------------
// conversions
// ---------------------------------------------------------------------------------------
__global__ void u8_to_f32_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, int64_t n, float mul, int use_div) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = (float)in[i];
    out[i] = use_div ? __fdiv_rn(v, mul) : v;
  }
}

__global__ void u8_to_f32_vec_kernel(const uint32_t *__restrict__ in, float4 *__restrict__ out, int64_t n4, float mul, int use_div) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t p = in[i];
    float4 v = make_float4((float)(p & 0xff), (float)((p >> 8) & 0xff), (float)((p >> 16) & 0xff), (float)(p >> 24));
    if (use_div) {
      v.x = __fdiv_rn(v.x, mul); v.y = __fdiv_rn(v.y, mul); v.z = __fdiv_rn(v.z, mul); v.w = __fdiv_rn(v.w, mul);
    }
    __stcs(out + i, v);
  }
}

int launch_u8_to_f32(ocrb_ctx *ctx, const uint8_t *in, int64_t n, float div, float *out) {
  if (n <= 0) return OCRB_OK;
  int use_div = div != 1.0f;
  bool aligned = ((uintptr_t)in % 4 == 0) && ((uintptr_t)out % 16 == 0);
  int64_t n4 = aligned ? n / 4 : 0;
  if (n4 > 0) {
    int64_t blocks = cdiv(n4, 256);
    int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    u8_to_f32_vec_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>((const uint32_t *)in, (float4 *)out, n4, div, use_div);
    OCRB_TRY(check_launch(ctx, "u8_to_f32_vec"));
  }
  int64_t rem = n - n4 * 4;
  if (rem > 0) {
    u8_to_f32_kernel<<<(unsigned)cdiv(rem, 256), 256, 0, ctx->stream>>>(in + n4 * 4, out + n4 * 4, rem, div, use_div);
    OCRB_TRY(check_launch(ctx, "u8_to_f32"));
  }
  return OCRB_OK;
}

// to_kind(Uint8) on a float tensor: C-style truncation toward zero, wrapping mod 256 like
// libtorch's static_cast chain float -> int64 -> uint8.
__global__ void f32_to_u8_kernel(const float *__restrict__ in, uint8_t *__restrict__ out, int64_t n, float scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = in[i] * scale;
    long long q = (long long)v;
    out[i] = (uint8_t)(q & 0xff);
  }
}

int launch_f32_to_u8(ocrb_ctx *ctx, const float *in, int64_t n, float scale, uint8_t *out) {
  if (n <= 0) return OCRB_OK;
  int64_t blocks = cdiv(n, 256);
  int64_t cap = (int64_t)ctx->sm_count * 16;
  if (blocks > cap) blocks = cap;
  f32_to_u8_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(in, out, n, scale);
  return check_launch(ctx, "f32_to_u8");
}

// fp32 [n][C] -> bf16 term planes [n][terms * C]: x = hi + mid (+ lo), each term the bf16 rounding of what is left
// (conv_tc.cuh: operands of the FP32-accuracy mode on the tensor cores)
__global__ void split_terms_kernel(const float *__restrict__ in, int64_t n, int C, int terms, __nv_bfloat16 *__restrict__ out) {
  const int64_t total = n * (C / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / (C / 4);
    const int c = (int)(i - row * (C / 4)) * 4;
    const float4 v = *reinterpret_cast<const float4 *>(in + row * C + c);
    float r[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 *o = out + row * (int64_t)(terms * C) + c;
    for (int t = 0; t < terms; ++t) {
      __nv_bfloat16 h[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        h[e] = __float2bfloat16_rn(r[e]);
        r[e] -= __bfloat162float(h[e]);
      }
      *reinterpret_cast<uint2 *>(o + (int64_t)t * C) = *reinterpret_cast<uint2 *>(h);
    }
  }
}

int launch_split_terms(ocrb_ctx *ctx, const float *in, int64_t n, int C, int terms, void *out) {
  if (n <= 0) return OCRB_OK;
  int64_t blocks = cdiv(n * (C / 4), 256);
  const int64_t cap = (int64_t)ctx->sm_count * 32;
  if (blocks > cap) blocks = cap;
  split_terms_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(in, n, C, terms, reinterpret_cast<__nv_bfloat16 *>(out));
  return check_launch(ctx, "split_terms");
}

// ---------------------------------------------------------------------------------------
// preprocess_image: Triangle resize (vertical pass, then horizontal), u8 intermediate with
// truncating stores, luma, zero pad.  One thread per output sample; the weight loop is the
// reference's scalar loop verbatim (same order, no FMA, f32 division by the weight sum).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float tri(float x) {
  float a = fabsf(x);
  return a < 1.0f ? 1.0f - a : 0.0f;
}

struct Taps {
  int left, right;
  float inputc2, sratio;
};

__device__ __forceinline__ Taps make_taps(int o, int n_in, int n_out) {
  float ratio = (float)n_in / (float)n_out;
  float sratio = ratio < 1.0f ? 1.0f : ratio;
  float support = 1.0f * sratio;
  float inputc = ((float)o + 0.5f) * ratio;
  long long left = (long long)floorf(inputc - support);
  if (left < 0) left = 0;
  if (left > n_in - 1) left = n_in - 1;
  long long right = (long long)ceilf(inputc + support);
  if (right < left + 1) right = left + 1;
  if (right > n_in) right = n_in;
  Taps t;
  t.left = (int)left;
  t.right = (int)right;
  t.inputc2 = inputc - 0.5f;
  t.sratio = sratio;
  return t;
}

// src [sh][sw][4] -> dst [rh][sw][4]
__global__ void resize_vertical_kernel(const uint8_t *__restrict__ src, int sw, int sh, int rh, uint8_t *__restrict__ dst) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)rh * sw;
  if (idx >= total) return;
  int x = (int)(idx % sw), o = (int)(idx / sw);
  Taps tp = make_taps(o, sh, rh);
  float sum = 0.0f, t0 = 0.0f, t1 = 0.0f, t2 = 0.0f, t3 = 0.0f;
  for (int i = tp.left; i < tp.right; ++i) {
    float w = tri(((float)i - tp.inputc2) / tp.sratio);
    sum += w;
    uchar4 p = *reinterpret_cast<const uchar4 *>(src + ((int64_t)i * sw + x) * 4);
    t0 += (float)p.x * w; t1 += (float)p.y * w; t2 += (float)p.z * w; t3 += (float)p.w * w;
  }
  t0 = t0 / sum; t1 = t1 / sum; t2 = t2 / sum; t3 = t3 / sum;
  uchar4 r;
  r.x = (uint8_t)fminf(fmaxf(t0, 0.0f), 255.0f);
  r.y = (uint8_t)fminf(fmaxf(t1, 0.0f), 255.0f);
  r.z = (uint8_t)fminf(fmaxf(t2, 0.0f), 255.0f);
  r.w = (uint8_t)fminf(fmaxf(t3, 0.0f), 255.0f);
  *reinterpret_cast<uchar4 *>(dst + ((int64_t)o * sw + x) * 4) = r;
}

// tmp [rh][sw][4] -> luma -> out [H][W] (zero padded).  `identity` skips the resample.
__global__ void resize_horizontal_luma_pad_kernel(const uint8_t *__restrict__ tmp, int sw, int rw, int rh, int W, int H,
                                                  int identity, uint8_t *__restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)W * H) return;
  int x = (int)(idx % W), y = (int)(idx / W);
  if (x >= rw || y >= rh) {
    out[idx] = 0;
    return;
  }
  float c0, c1, c2;
  if (identity) {
    uchar4 p = *reinterpret_cast<const uchar4 *>(tmp + ((int64_t)y * sw + x) * 4);
    c0 = (float)p.x; c1 = (float)p.y; c2 = (float)p.z;
  } else {
    Taps tp = make_taps(x, sw, rw);
    float sum = 0.0f, t0 = 0.0f, t1 = 0.0f, t2 = 0.0f;
    for (int i = tp.left; i < tp.right; ++i) {
      float w = tri(((float)i - tp.inputc2) / tp.sratio);
      sum += w;
      uchar4 p = *reinterpret_cast<const uchar4 *>(tmp + ((int64_t)y * sw + i) * 4);
      t0 += (float)p.x * w; t1 += (float)p.y * w; t2 += (float)p.z * w;
    }
    c0 = (float)(uint8_t)fminf(fmaxf(t0 / sum, 0.0f), 255.0f);
    c1 = (float)(uint8_t)fminf(fmaxf(t1 / sum, 0.0f), 255.0f);
    c2 = (float)(uint8_t)fminf(fmaxf(t2 / sum, 0.0f), 255.0f);
  }
  float l = 0.2126f * c0 + 0.7152f * c1 + 0.0722f * c2;  // -fmad=false: ((a+b)+c) in f32
  out[idx] = (uint8_t)l;
}

// ---------------------------------------------------------------------------------------
// The same for a BATCH, fused: one CTA per 16 x 128 output tile runs the vertical pass for the
// source-column span its outputs need into shared memory (u8, truncated: the reference's
// intermediate image, never written to HBM) and the horizontal pass + luma + pad from there.
// HBM traffic = the RGBA source once (neighbouring tiles re-read their overlap from L2) + the
// grey output once.  Per-image geometry comes from a small descriptor array.
// ---------------------------------------------------------------------------------------
struct PreImage {
  int64_t src_off;  // byte offset of the image's RGBA8 pixels in the packed source buffer
  int sw, sh, rw, rh;
};
constexpr int PB_TH = 16, PB_TW = 128, PB_THREADS = 256;  // 32-row tiles measured 40 % slower (occupancy)
constexpr int PB_SPAN = 1400;  // most source columns a tile may span (shared memory: 16 x 1400 x 4 B = 87.5 KB)
constexpr int PB_TAPS = 32;    // most filter taps per output sample kept in shared memory (down-scaling up to ~15x)

// filter taps of one output sample, evaluated once per tile row / tile column instead of once per pixel; the weights
// and their sum are the reference's own f32 values (same expressions, same summation order)
struct PbTaps { int left, n; float sum; };

__global__ void __launch_bounds__(PB_THREADS) preprocess_batch_kernel(const uint8_t *__restrict__ rgba, const PreImage *__restrict__ imgs, int W, int H,
                                                                      int span_cap, uint8_t *__restrict__ out, int *__restrict__ overflow) {
  extern __shared__ uchar4 s_tmp[];  // [PB_TH][span_cap]
  __shared__ PbTaps s_vt[PB_TH], s_ht[PB_TW];
  __shared__ float s_vw[PB_TH][PB_TAPS], s_hw[PB_TW][PB_TAPS];
  const PreImage im = imgs[blockIdx.z];
  const int x0 = blockIdx.x * PB_TW, y0 = blockIdx.y * PB_TH;
  uint8_t *dst = out + (int64_t)blockIdx.z * H * W;
  const uchar4 *src = reinterpret_cast<const uchar4 *>(rgba + im.src_off);
  const int x1 = min(x0 + PB_TW, W), y1 = min(y0 + PB_TH, H);
  const int tw = x1 - x0;
  // rows / columns outside the resized image: zero padding (image_ops.rs:204-214)
  if (x0 >= im.rw || y0 >= im.rh) {
    for (int i = threadIdx.x; i < (y1 - y0) * tw; i += PB_THREADS) dst[(int64_t)(y0 + i / tw) * W + x0 + i % tw] = 0;
    return;
  }
  const bool identity = im.rw == im.sw && im.rh == im.sh;
  const int xe = min(x1, im.rw), ye = min(y1, im.rh);  // valid outputs of this tile
  int c0 = x0, c1 = xe;                                // source columns needed
  if (!identity) {
    c0 = make_taps(x0, im.sw, im.rw).left;
    c1 = make_taps(xe - 1, im.sw, im.rw).right;
    // taps of the tile's rows and columns
    for (int i = threadIdx.x; i < PB_TH + PB_TW; i += PB_THREADS) {
      const bool vert = i < PB_TH;
      const int o = vert ? y0 + i : x0 + (i - PB_TH);
      if (o >= (vert ? ye : xe)) continue;
      const Taps tp = vert ? make_taps(o, im.sh, im.rh) : make_taps(o, im.sw, im.rw);
      PbTaps pt;
      pt.left = tp.left;
      pt.n = tp.right - tp.left;
      float sum = 0.0f;
      float *wv = vert ? s_vw[i] : s_hw[i - PB_TH];
      for (int k = 0; k < pt.n && k < PB_TAPS; ++k) {
        const float w = tri(((float)(tp.left + k) - tp.inputc2) / tp.sratio);
        wv[k] = w;
        sum += w;
      }
      pt.sum = sum;
      if (pt.n > PB_TAPS) atomicExch(overflow, 1);
      if (vert) s_vt[i] = pt; else s_ht[i - PB_TH] = pt;
    }
  }
  const int span = c1 - c0;
  if (span > span_cap) {  // cannot happen when the host sized span_cap from the batch; kept as a guard
    if (threadIdx.x == 0) atomicExch(overflow, 1);
    return;
  }
  __syncthreads();
  // ---- vertical pass: tmp[r][c] for output rows y0 .. ye-1, source columns c0 .. c1-1 (a warp walks one row: coalesced, no division)
  for (int r = threadIdx.x >> 5; r < ye - y0; r += PB_THREADS / 32)
  for (int c = threadIdx.x & 31; c < span; c += 32) {
    uchar4 v;
    if (identity) {
      v = src[(int64_t)(y0 + r) * im.sw + c0 + c];
    } else {
      const PbTaps pt = s_vt[r];
      const uchar4 *col = src + (int64_t)pt.left * im.sw + c0 + c;
      float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f, t3 = 0.0f;
      for (int k = 0; k < pt.n; ++k) {
        const float w = s_vw[r][k];
        const uchar4 p = col[(int64_t)k * im.sw];
        t0 += (float)p.x * w; t1 += (float)p.y * w; t2 += (float)p.z * w; t3 += (float)p.w * w;
      }
      t0 = t0 / pt.sum; t1 = t1 / pt.sum; t2 = t2 / pt.sum; t3 = t3 / pt.sum;
      v.x = (uint8_t)fminf(fmaxf(t0, 0.0f), 255.0f);
      v.y = (uint8_t)fminf(fmaxf(t1, 0.0f), 255.0f);
      v.z = (uint8_t)fminf(fmaxf(t2, 0.0f), 255.0f);
      v.w = (uint8_t)fminf(fmaxf(t3, 0.0f), 255.0f);
    }
    s_tmp[r * span + c] = v;
  }
  __syncthreads();
  // ---- horizontal pass + luma + pad: a thread owns 4 neighbouring output columns of one row (one 32-bit store)
  for (int i = threadIdx.x; i < (y1 - y0) * (PB_TW / 4); i += PB_THREADS) {
    const int r = i / (PB_TW / 4), xq = (i - r * (PB_TW / 4)) * 4, y = y0 + r;
    if (x0 + xq >= x1) continue;
    uint32_t packed = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int x = x0 + xq + e;
      uint8_t g = 0;
      if (x < im.rw && y < im.rh) {
        float a0, a1, a2;
        if (identity) {
          const uchar4 p = s_tmp[r * span + (x - c0)];
          a0 = (float)p.x; a1 = (float)p.y; a2 = (float)p.z;
        } else {
          const PbTaps pt = s_ht[xq + e];
          const uchar4 *row = s_tmp + r * span + (pt.left - c0);
          float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f;
          for (int k = 0; k < pt.n; ++k) {
            const float w = s_hw[xq + e][k];
            const uchar4 p = row[k];
            t0 += (float)p.x * w; t1 += (float)p.y * w; t2 += (float)p.z * w;
          }
          a0 = (float)(uint8_t)fminf(fmaxf(t0 / pt.sum, 0.0f), 255.0f);
          a1 = (float)(uint8_t)fminf(fmaxf(t1 / pt.sum, 0.0f), 255.0f);
          a2 = (float)(uint8_t)fminf(fmaxf(t2 / pt.sum, 0.0f), 255.0f);
        }
        const float l = 0.2126f * a0 + 0.7152f * a1 + 0.0722f * a2;  // -fmad=false: ((a+b)+c) in f32
        g = (uint8_t)l;
      }
      packed |= (uint32_t)g << (8 * e);
    }
    if (x0 + xq + 4 <= x1 && (W & 3) == 0) {
      *reinterpret_cast<uint32_t *>(dst + (int64_t)y * W + x0 + xq) = packed;
    } else {
      for (int e = 0; e < 4 && x0 + xq + e < x1; ++e) dst[(int64_t)y * W + x0 + xq + e] = (uint8_t)(packed >> (8 * e));
    }
  }
}

// every image of the batch already has the target size: luma + pad only, pure streaming (16 source bytes -> 4 grey bytes per thread step)
__global__ void preprocess_batch_identity_kernel(const uint8_t *__restrict__ rgba, const PreImage *__restrict__ imgs, int W, int H,
                                                 uint8_t *__restrict__ out) {
  const PreImage im = imgs[blockIdx.y];
  const int qw = W / 4, quads = H * qw;
  uint8_t *dst = out + (int64_t)blockIdx.y * H * W;
  const bool aligned = ((reinterpret_cast<uintptr_t>(rgba) + im.src_off) & 15) == 0 && (im.sw & 3) == 0;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += gridDim.x * blockDim.x) {
    const int y = q / qw, x = (q - y * qw) * 4;
    uint32_t packed = 0;
    if (y < im.rh && x < im.rw) {
      const uint8_t *src = rgba + im.src_off + ((int64_t)y * im.sw + x) * 4;
      uint32_t px[4] = {0u, 0u, 0u, 0u};
      if (aligned && x + 4 <= im.rw) {
        const uint4 v = *reinterpret_cast<const uint4 *>(src);  // four RGBA pixels in one 16-byte load
        px[0] = v.x; px[1] = v.y; px[2] = v.z; px[3] = v.w;
      } else {
        for (int e = 0; e < 4 && x + e < im.rw; ++e) px[e] = *reinterpret_cast<const uint32_t *>(src + 4 * e);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (x + e >= im.rw) break;
        const float l = 0.2126f * (float)(px[e] & 0xffu) + 0.7152f * (float)((px[e] >> 8) & 0xffu) + 0.0722f * (float)((px[e] >> 16) & 0xffu);
        packed |= (uint32_t)(uint8_t)l << (8 * e);
      }
    }
    *reinterpret_cast<uint32_t *>(dst + (int64_t)y * W + x) = packed;
  }
}

int launch_preprocess_batch_identity(ocrb_ctx *ctx, const uint8_t *rgba_dev, const void *imgs_dev, int n, int W, int H, uint8_t *out_dev) {
  const int quads = H * (W / 4);
  int bx = (int)cdiv(quads, 256 * 4);  // four 16-byte loads in flight per thread
  if (bx < 1) bx = 1;
  preprocess_batch_identity_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, ctx->stream>>>(rgba_dev, reinterpret_cast<const PreImage *>(imgs_dev), W, H, out_dev);
  return check_launch(ctx, "preprocess_batch_identity");
}

// span_cap: source columns a tile may need for this batch (host-computed from the largest down-scaling factor); 0 = too wide
int launch_preprocess_batch(ocrb_ctx *ctx, const uint8_t *rgba_dev, const void *imgs_dev, int n, int W, int H, int span_cap, uint8_t *out_dev,
                            int *overflow_dev) {
  if (n <= 0) return OCRB_OK;
  const int smem = PB_TH * span_cap * 4;
  OCRB_TRY(ensure_dyn_smem(ctx, preprocess_batch_kernel, PB_TH * PB_SPAN * 4));
  dim3 grid((unsigned)cdiv(W, PB_TW), (unsigned)cdiv(H, PB_TH), (unsigned)n);
  preprocess_batch_kernel<<<grid, PB_THREADS, smem, ctx->stream>>>(rgba_dev, reinterpret_cast<const PreImage *>(imgs_dev), W, H, span_cap, out_dev,
                                                                   overflow_dev);
  return check_launch(ctx, "preprocess_batch");
}
int preprocess_batch_span_limit() { return PB_SPAN; }
int preprocess_batch_tile_width() { return PB_TW; }

int launch_preprocess(ocrb_ctx *ctx, const uint8_t *rgba_dev, int sw, int sh, int rw, int rh, int W, int H,
                      uint8_t *tmp_dev, uint8_t *out_dev) {
  int identity = (rw == sw && rh == sh);
  const uint8_t *hsrc = rgba_dev;
  if (!identity) {
    int64_t total = (int64_t)rh * sw;
    resize_vertical_kernel<<<(unsigned)cdiv(total, 256), 256, 0, ctx->stream>>>(rgba_dev, sw, sh, rh, tmp_dev);
    OCRB_TRY(check_launch(ctx, "resize_vertical"));
    hsrc = tmp_dev;
  }
  resize_horizontal_luma_pad_kernel<<<(unsigned)cdiv((int64_t)W * H, 256), 256, 0, ctx->stream>>>(hsrc, sw, rw, rh, W, H,
                                                                                                identity, out_dev);
  return check_launch(ctx, "resize_horizontal_luma_pad");
}

}  // namespace ocrb
