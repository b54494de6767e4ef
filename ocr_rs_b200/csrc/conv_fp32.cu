// FP32 mode of the detector (north_star: probability maps within 1e-4 of the reference's
// libtorch fp32 path).  CUDA-core kernels, NHWC fp32 activations, fp32 accumulation.
// This mode is the accuracy reference on the device and the cross-check for the BF16
// tcgen05 engine (conv_tc.cu); it is not the throughput path.
//
// reference ops: model.rs:4-12 (conv2d, no bias), :14-28 (conv_transpose2d k2 s2 + bias),
// batch_norm eval folded to scale/shift, relu, max_pool2d(3,2,1), upsample_nearest2d, cat.
#include "common.cuh"

namespace ocrb {

// ---------------------------------------------------------------------------------------
// stem: conv 7x7 s2 p3 (1 -> 64) + BN + ReLU.  in [B][H][W] f32, out [B][H/2][W/2][64]
// One thread per (pixel, 4 channels); weights [49][64] in shared memory.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stem_conv_fp32_kernel(const float *__restrict__ in, int B, int H, int W,
                                                             const float *__restrict__ w /*[49][64]*/,
                                                             const float *__restrict__ scale, const float *__restrict__ shift,
                                                             float *__restrict__ out) {
  __shared__ float ws[49 * 64];
  for (int i = threadIdx.x; i < 49 * 64; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int Ho = H / 2, Wo = W / 2;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)B * Ho * Wo * 16;
  if (idx >= total) return;
  int cg = (int)(idx % 16);
  int64_t p = idx / 16;
  int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho);
  int64_t b = p / ((int64_t)Wo * Ho);
  const float *img = in + b * H * W;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = 0; r < 7; ++r) {
    int iy = oy * 2 + r - 3;
    if (iy < 0 || iy >= H) continue;
    for (int s = 0; s < 7; ++s) {
      int ix = ox * 2 + s - 3;
      if (ix < 0 || ix >= W) continue;
      float v = img[(int64_t)iy * W + ix];
      const float *wp = ws + (r * 7 + s) * 64 + cg * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(v, wp[j], acc[j]);
    }
  }
  float4 o;
  float *op = &o.x;
#pragma unroll
  for (int j = 0; j < 4; ++j) op[j] = fmaxf(fmaf(acc[j], scale[cg * 4 + j], shift[cg * 4 + j]), 0.f);
  *reinterpret_cast<float4 *>(out + p * 64 + cg * 4) = o;
}

// max_pool2d k3 s2 p1 (ceil_mode false), NHWC, C % 4 == 0
__global__ void maxpool3x3s2_fp32_kernel(const float *__restrict__ in, int B, int H, int W, int C, float *__restrict__ out) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1, C4 = C / 4;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)B * Ho * Wo * C4;
  if (idx >= total) return;
  int c4 = (int)(idx % C4);
  int64_t p = idx / C4;
  int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho);
  int64_t b = p / ((int64_t)Wo * Ho);
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (int r = 0; r < 3; ++r) {
    int iy = oy * 2 + r - 1;
    if (iy < 0 || iy >= H) continue;
    for (int s = 0; s < 3; ++s) {
      int ix = ox * 2 + s - 1;
      if (ix < 0 || ix >= W) continue;
      float4 v = *reinterpret_cast<const float4 *>(in + ((b * H + iy) * W + ix) * C + c4 * 4);
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  }
  *reinterpret_cast<float4 *>(out + p * C + c4 * 4) = m;
}

// ---------------------------------------------------------------------------------------
// generic conv (R x S in {1,3}, stride {1,2}, pad) as a shared-memory tiled implicit GEMM:
// CTA tile 64 pixels x 64 output channels, K chunks of 16 input channels per tap,
// 256 threads x (4 pixels x 4 channels).  Epilogue: y = acc*scale + shift (+ residual) (ReLU).
// in [B][H][W][Cin], w [R*S][Cin][Cout], out [B][Ho][Wo][Cout].  Cin % 16 == 0, Cout % 64 == 0.
// ---------------------------------------------------------------------------------------
struct ConvFp32Params {
  const float *in, *w, *scale, *shift, *residual;
  float *out;
  int B, H, W, Cin, Ho, Wo, Cout, R, S, stride, pad, relu;
};

__global__ void __launch_bounds__(256) conv_fp32_kernel(ConvFp32Params p) {
  __shared__ __align__(16) float As[16][64 + 4];
  __shared__ __align__(16) float Bs[16][64];
  const int t = threadIdx.x;
  const int64_t M = (int64_t)p.B * p.Ho * p.Wo;
  const int64_t m0 = (int64_t)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  // A-load role: pixel lm, channel group kg
  const int lm = t >> 2, kg = t & 3;
  int64_t pm = m0 + lm;
  bool pm_ok = pm < M;
  int ox = 0, oy = 0;
  int64_t b = 0;
  if (pm_ok) {
    ox = (int)(pm % p.Wo);
    oy = (int)((pm / p.Wo) % p.Ho);
    b = pm / ((int64_t)p.Wo * p.Ho);
  }
  // B-load role
  const int bk = t >> 4, bn4 = t & 15;
  // compute role
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int r = 0; r < p.R; ++r) {
    for (int s = 0; s < p.S; ++s) {
      int iy = oy * p.stride + r - p.pad, ix = ox * p.stride + s - p.pad;
      bool ok = pm_ok && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
      const float *ap = p.in + ((b * p.H + iy) * p.W + ix) * p.Cin + kg * 4;
      const float *wp = p.w + ((int64_t)(r * p.S + s) * p.Cin + bk) * p.Cout + n0 + bn4 * 4;
      for (int c0 = 0; c0 < p.Cin; c0 += 16) {
        float4 av = ok ? *reinterpret_cast<const float4 *>(ap + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 bv = *reinterpret_cast<const float4 *>(wp + (int64_t)c0 * p.Cout);
        __syncthreads();
        As[kg * 4 + 0][lm] = av.x; As[kg * 4 + 1][lm] = av.y; As[kg * 4 + 2][lm] = av.z; As[kg * 4 + 3][lm] = av.w;
        *reinterpret_cast<float4 *>(&Bs[bk][bn4 * 4]) = bv;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
          float4 bb = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
          const float av4[4] = {a.x, a.y, a.z, a.w}, bv4[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av4[i], bv4[j], acc[i][j]);
        }
      }
    }
  }
  const int n = n0 + tx * 4;
  float4 sc = *reinterpret_cast<const float4 *>(p.scale + n), sh = *reinterpret_cast<const float4 *>(p.shift + n);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    float4 o = make_float4(fmaf(acc[i][0], sc.x, sh.x), fmaf(acc[i][1], sc.y, sh.y), fmaf(acc[i][2], sc.z, sh.z),
                           fmaf(acc[i][3], sc.w, sh.w));
    if (p.residual) {
      float4 rv = *reinterpret_cast<const float4 *>(p.residual + m * p.Cout + n);
      o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
    }
    if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    *reinterpret_cast<float4 *>(p.out + m * p.Cout + n) = o;
  }
}

// out[b][y][x][c] = up2(a)[...] + bsrc[...]; a [B][H/2][W/2][C], bsrc/out [B][H][W][C]
__global__ void upsample2_add_fp32_kernel(const float *__restrict__ a, const float *__restrict__ bsrc, int B, int H, int W, int C,
                                          float *__restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int C4 = C / 4;
  int64_t total = (int64_t)B * H * W * C4;
  if (idx >= total) return;
  int c4 = (int)(idx % C4);
  int64_t p = idx / C4;
  int x = (int)(p % W), y = (int)((p / W) % H);
  int64_t b = p / ((int64_t)W * H);
  float4 u = *reinterpret_cast<const float4 *>(a + ((b * (H / 2) + y / 2) * (W / 2) + x / 2) * C + c4 * 4);
  float4 v = *reinterpret_cast<const float4 *>(bsrc + p * C + c4 * 4);
  *reinterpret_cast<float4 *>(out + p * C + c4 * 4) = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
}

// nearest upsample by f into a channel slice of the concat buffer:
// src [B][H/f][W/f][C] -> dst[b][y][x][c_off + c], dst row pitch ldc
__global__ void upsample_concat_fp32_kernel(const float *__restrict__ src, int B, int H, int W, int C, int f, int c_off, int ldc,
                                            float *__restrict__ dst) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int C4 = C / 4;
  int64_t total = (int64_t)B * H * W * C4;
  if (idx >= total) return;
  int c4 = (int)(idx % C4);
  int64_t p = idx / C4;
  int x = (int)(p % W), y = (int)((p / W) % H);
  int64_t b = p / ((int64_t)W * H);
  float4 v = *reinterpret_cast<const float4 *>(src + ((b * (H / f) + y / f) * (W / f) + x / f) * C + c4 * 4);
  *reinterpret_cast<float4 *>(dst + p * ldc + c_off + c4 * 4) = v;
}

// conv_transpose2d k2 s2 (Cin -> Cout) + bias folded into shift + BN + ReLU
// in [B][H][W][Cin], w [2*2][Cin][Cout], out [B][2H][2W][Cout]
__global__ void __launch_bounds__(256) convt2x2_fp32_kernel(const float *__restrict__ in, int B, int H, int W, int Cin, int Cout,
                                                            const float *__restrict__ w, const float *__restrict__ scale,
                                                            const float *__restrict__ shift, int relu, float *__restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int Ho = 2 * H, Wo = 2 * W, C4 = Cout / 4;
  int64_t total = (int64_t)B * Ho * Wo * C4;
  if (idx >= total) return;
  int c4 = (int)(idx % C4);
  int64_t p = idx / C4;
  int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho);
  int64_t b = p / ((int64_t)Wo * Ho);
  int tap = (oy & 1) * 2 + (ox & 1);
  const float *ip = in + ((b * H + oy / 2) * W + ox / 2) * Cin;
  const float *wp = w + (int64_t)tap * Cin * Cout + c4 * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int ci = 0; ci < Cin; ++ci) {
    float v = ip[ci];
    float4 ww = *reinterpret_cast<const float4 *>(wp + (int64_t)ci * Cout);
    acc.x = fmaf(v, ww.x, acc.x); acc.y = fmaf(v, ww.y, acc.y); acc.z = fmaf(v, ww.z, acc.z); acc.w = fmaf(v, ww.w, acc.w);
  }
  float4 sc = *reinterpret_cast<const float4 *>(scale + c4 * 4), sh = *reinterpret_cast<const float4 *>(shift + c4 * 4);
  float4 o = make_float4(fmaf(acc.x, sc.x, sh.x), fmaf(acc.y, sc.y, sh.y), fmaf(acc.z, sc.z, sh.z), fmaf(acc.w, sc.w, sh.w));
  if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
  *reinterpret_cast<float4 *>(out + p * Cout + c4 * 4) = o;
}

// final conv_transpose2d k2 s2 (Cin -> 1) + bias + sigmoid.  in [B][H][W][Cin], w [4][Cin],
// out [B][2H][2W] f32
// tap_major: `in` is [B][H / 2][W / 2][4 taps][Cin] (conv-transpose 1 computed as a 1x1 convolution 64 -> 4 * 64) instead of [B][H][W][Cin]
__global__ void convt2x2_sigmoid_fp32_kernel(const float *__restrict__ in, int B, int H, int W, int Cin,
                                             const float *__restrict__ w, float bias, float *__restrict__ out, int tap_major) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int Ho = 2 * H, Wo = 2 * W;
  int64_t total = (int64_t)B * Ho * Wo;
  if (idx >= total) return;
  int ox = (int)(idx % Wo), oy = (int)((idx / Wo) % Ho);
  int64_t b = idx / ((int64_t)Wo * Ho);
  int tap = (oy & 1) * 2 + (ox & 1);
  const int iy = oy / 2, ix = ox / 2;  // pixel of the H x W input map
  const float *ip = tap_major ? in + ((((b * (H / 2) + iy / 2) * (W / 2) + ix / 2) * 4 + (iy & 1) * 2 + (ix & 1)) * (int64_t)Cin)
                              : in + ((b * H + iy) * W + ix) * Cin;
  const float *wp = w + tap * Cin;
  float acc = 0.f;
  for (int ci = 0; ci < Cin; ci += 4) {
    float4 v = *reinterpret_cast<const float4 *>(ip + ci);
    float4 ww = *reinterpret_cast<const float4 *>(wp + ci);
    acc = fmaf(v.x, ww.x, acc); acc = fmaf(v.y, ww.y, acc); acc = fmaf(v.z, ww.z, acc); acc = fmaf(v.w, ww.w, acc);
  }
  float z = acc + bias;
  out[idx] = 1.0f / (1.0f + expf(-z));
}

// NHWC f32 -> NCHW f32 (test taps)
__global__ void nhwc_to_nchw_fp32_kernel(const float *__restrict__ in, int B, int H, int W, int C, int ldc, float *__restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)B * C * H * W;
  if (idx >= total) return;
  int x = (int)(idx % W), y = (int)((idx / W) % H), c = (int)((idx / ((int64_t)W * H)) % C);
  int64_t b = idx / ((int64_t)W * H * C);
  out[idx] = in[((b * H + y) * W + x) * ldc + c];
}

// ---- launchers ---------------------------------------------------------------------------
int launch_stem_fp32(ocrb_ctx *ctx, const float *in, int B, int H, int W, const float *w, const float *scale, const float *shift,
                     float *out) {
  int64_t total = (int64_t)B * (H / 2) * (W / 2) * 16;
  stem_conv_fp32_kernel<<<(unsigned)cdiv(total, 256), 256, 0, ctx->stream>>>(in, B, H, W, w, scale, shift, out);
  return check_launch(ctx, "stem_conv_fp32");
}
int launch_maxpool_fp32(ocrb_ctx *ctx, const float *in, int B, int H, int W, int C, float *out) {
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  int64_t total = (int64_t)B * Ho * Wo * (C / 4);
  maxpool3x3s2_fp32_kernel<<<(unsigned)cdiv(total, 256), 256, 0, ctx->stream>>>(in, B, H, W, C, out);
  return check_launch(ctx, "maxpool_fp32");
}
int launch_conv_fp32(ocrb_ctx *ctx, const float *in, int B, int H, int W, int Cin, const float *w, int Cout, int R, int stride,
                     int pad, const float *scale, const float *shift, const float *residual, int relu, float *out) {
  ConvFp32Params p;
  p.in = in; p.w = w; p.scale = scale; p.shift = shift; p.residual = residual; p.out = out;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.R = R; p.S = R; p.stride = stride; p.pad = pad; p.relu = relu;
  p.Ho = (H + 2 * pad - R) / stride + 1;
  p.Wo = (W + 2 * pad - R) / stride + 1;
  if (Cin % 16 != 0 || Cout % 64 != 0) { set_error("conv_fp32: Cin %% 16 / Cout %% 64 violated (%d, %d)", Cin, Cout); return OCRB_ERR_INVALID; }
  int64_t M = (int64_t)B * p.Ho * p.Wo;
  dim3 grid((unsigned)cdiv(M, 64), (unsigned)(Cout / 64));
  conv_fp32_kernel<<<grid, 256, 0, ctx->stream>>>(p);
  return check_launch(ctx, "conv_fp32");
}
int launch_upsample2_add_fp32(ocrb_ctx *ctx, const float *a, const float *b, int B, int H, int W, int C, float *out) {
  int64_t total = (int64_t)B * H * W * (C / 4);
  upsample2_add_fp32_kernel<<<(unsigned)cdiv(total, 256), 256, 0, ctx->stream>>>(a, b, B, H, W, C, out);
  return check_launch(ctx, "upsample2_add_fp32");
}
int launch_upsample_concat_fp32(ocrb_ctx *ctx, const float *src, int B, int H, int W, int C, int f, int c_off, int ldc, float *dst) {
  int64_t total = (int64_t)B * H * W * (C / 4);
  upsample_concat_fp32_kernel<<<(unsigned)cdiv(total, 256), 256, 0, ctx->stream>>>(src, B, H, W, C, f, c_off, ldc, dst);
  return check_launch(ctx, "upsample_concat_fp32");
}
int launch_convt2x2_fp32(ocrb_ctx *ctx, const float *in, int B, int H, int W, int Cin, int Cout, const float *w, const float *scale,
                         const float *shift, int relu, float *out) {
  int64_t total = (int64_t)B * 2 * H * 2 * W * (Cout / 4);
  convt2x2_fp32_kernel<<<(unsigned)cdiv(total, 256), 256, 0, ctx->stream>>>(in, B, H, W, Cin, Cout, w, scale, shift, relu, out);
  return check_launch(ctx, "convt2x2_fp32");
}
int launch_convt2x2_sigmoid_fp32(ocrb_ctx *ctx, const float *in, int B, int H, int W, int Cin, const float *w, float bias, float *out, int tap_major) {
  int64_t total = (int64_t)B * 2 * H * 2 * W;
  convt2x2_sigmoid_fp32_kernel<<<(unsigned)cdiv(total, 256), 256, 0, ctx->stream>>>(in, B, H, W, Cin, w, bias, out, tap_major);
  return check_launch(ctx, "convt2x2_sigmoid_fp32");
}
int launch_nhwc_to_nchw_fp32(ocrb_ctx *ctx, const float *in, int B, int H, int W, int C, int ldc, float *out) {
  int64_t total = (int64_t)B * C * H * W;
  nhwc_to_nchw_fp32_kernel<<<(unsigned)cdiv(total, 256), 256, 0, ctx->stream>>>(in, B, H, W, C, ldc, out);
  return check_launch(ctx, "nhwc_to_nchw_fp32");
}

}  // namespace ocrb
