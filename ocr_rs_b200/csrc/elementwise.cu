// HBM-bound element-wise kernels of the path: binarize, u8<->f32 conversions, and the
// preprocess_image resize/luma/pad.  Compiled with -fmad=false: the resize arithmetic must
// round exactly like the reference's scalar f32 code (image 0.23.11).
//
// reference: metrics.rs:129-131 (binarize), image_ops.rs:350-381 (conversions),
//            image_ops.rs:73-85 (/255), image_ops.rs:188-220 (preprocess_image)
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
