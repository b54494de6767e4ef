// text_detection::model::resnet18 (model.rs:65-156) as a device-resident graph:
// ResNet-18 (1-channel stem) -> 1x1 laterals -> non-cascaded FPN -> 3x3 "out"
// convs -> x8/x4/x2 nearest upsample + concat -> DB probability head -> sigmoid.
//
// Two arithmetic modes behind one entry point (north_star tolerances):
//   OCRB_MODE_FP32  CUDA-core fp32 kernels (conv_fp32.cu), 1e-4 class
//   OCRB_MODE_BF16  tcgen05/TMEM implicit-GEMM engine (conv_tc.cu), bf16 operands, fp32
//                   accumulation, batch-norm applied in the fp32 epilogue, 1e-2 class
// Weights arrive as the reference's VarStore tensors (OIHW fp32).
#include <cuda_bf16.h>

#include <map>
#include <string>

#include "common.cuh"
#include "conv_tc.cuh"

namespace ocrb {

// conv_fp32.cu
int launch_stem_fp32(ocrb_ctx *, const float *, int, int, int, const float *, const float *, const float *, float *);
int launch_maxpool_fp32(ocrb_ctx *, const float *, int, int, int, int, float *);
int launch_split_terms(ocrb_ctx *, const float *, int64_t, int, int, void *);
int launch_conv_fp32(ocrb_ctx *, const float *, int, int, int, int, const float *, int, int, int, int, const float *,
                     const float *, const float *, int, float *);
int launch_upsample2_add_fp32(ocrb_ctx *, const float *, const float *, int, int, int, int, float *);
int launch_upsample_concat_fp32(ocrb_ctx *, const float *, int, int, int, int, int, int, int, float *);
int launch_convt2x2_fp32(ocrb_ctx *, const float *, int, int, int, int, int, const float *, const float *, const float *, int, float *);
int launch_convt2x2_sigmoid_fp32(ocrb_ctx *, const float *, int, int, int, int, const float *, float, float *, int);
int launch_nhwc_to_nchw_fp32(ocrb_ctx *, const float *, int, int, int, int, int, float *);
int launch_u8_to_f32(ocrb_ctx *, const uint8_t *, int64_t, float, float *);
// stem_tc.cu
int launch_stem_tc(ocrb_ctx *, const void *, int, int, int, int, const float *, const float *, const float *, __nv_bfloat16 *, int *);
int launch_stem_tc2(ocrb_ctx *, const void *, int, int, int, int, const float *, const float *, const float *, __nv_bfloat16 *, int *);
int launch_stem_tc3(ocrb_ctx *, const void *, int, int, int, int, const float *, const float *, const float *, __nv_bfloat16 *, int *);

constexpr float BN_EPS = 1e-5f;  // tch nn::BatchNormConfig default

// ---------------------------------------------------------------------------------------
// BF16-mode stem: conv 7x7 s2 p3 (1 -> 64) + BN + ReLU + max_pool 3x3 s2 p1, fused; u8 or f32
// grey levels in, NHWC bf16 [B][H/4][W/4][64] out.  CUDA cores (K = 49 is too thin for a UMMA
// tile without an im2col staging pass).  CTA = 8 x 16 pooled pixels.
// ---------------------------------------------------------------------------------------
constexpr int ST_PH = 8, ST_PW = 16;                       // pooled tile
constexpr int ST_CH = 2 * ST_PH + 1, ST_CW = 2 * ST_PW + 1;  // conv tile 17 x 33
constexpr int ST_IH = 2 * ST_CH + 5, ST_IW = 2 * ST_CW + 5;  // input patch 39 x 71
constexpr int ST_THREADS = 288;

template <class TIn>
__global__ void __launch_bounds__(ST_THREADS) stem_fused_bf16_kernel(const TIn *__restrict__ in, int B, int H, int W,
                                                                     const float *__restrict__ w /*[49][64]*/,
                                                                     const float *__restrict__ scale, const float *__restrict__ shift,
                                                                     __nv_bfloat16 *__restrict__ out) {
  __shared__ float s_in[ST_IH * ST_IW];
  __shared__ __align__(16) float s_w[49 * 64];
  __shared__ __nv_bfloat16 s_conv[ST_CH * ST_CW * 16];
  const int Hc = H / 2, Wc = W / 2, Hp = H / 4, Wp = W / 4;
  const int tiles_x = (Wp + ST_PW - 1) / ST_PW, tiles_y = (Hp + ST_PH - 1) / ST_PH;
  const int b = blockIdx.x / (tiles_x * tiles_y), t = blockIdx.x % (tiles_x * tiles_y);
  const int py0 = (t / tiles_x) * ST_PH, px0 = (t % tiles_x) * ST_PW;
  const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;  // conv-grid origin of the tile (pool pad 1)
  const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;  // input origin (conv pad 3)
  const TIn *img = in + (int64_t)b * H * W;
  for (int i = threadIdx.x; i < ST_IH * ST_IW; i += ST_THREADS) {
    int yy = iy0 + i / ST_IW, xx = ix0 + i % ST_IW;
    s_in[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (float)img[(int64_t)yy * W + xx] : 0.0f;
  }
  for (int i = threadIdx.x; i < 49 * 64; i += ST_THREADS) s_w[i] = w[i];
  __syncthreads();
  for (int cg = 0; cg < 4; ++cg) {  // 16 output channels at a time
    for (int pos = threadIdx.x; pos < ST_CH * ST_CW; pos += ST_THREADS) {
      const int cy = pos / ST_CW, cx = pos % ST_CW;
      const bool in_grid = (cy0 + cy) >= 0 && (cy0 + cy) < Hc && (cx0 + cx) >= 0 && (cx0 + cx) < Wc;
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
      const float *ip = s_in + (2 * cy) * ST_IW + 2 * cx;
#pragma unroll 1
      for (int r = 0; r < 7; ++r) {
#pragma unroll
        for (int s = 0; s < 7; ++s) {
          const float v = ip[r * ST_IW + s];
          const float4 *wp = reinterpret_cast<const float4 *>(s_w + (r * 7 + s) * 64 + cg * 16);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 ww = wp[q];
            acc[4 * q + 0] = fmaf(v, ww.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(v, ww.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v, ww.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(v, ww.w, acc[4 * q + 3]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int c = cg * 16 + j;
        // positions outside the conv grid are max-pool padding (-inf); ReLU output >= 0, so
        // 0 would also be neutral only if a valid element exists -> use a large negative
        float y = in_grid ? fmaxf(fmaf(acc[j], scale[c], shift[c]), 0.0f) : -3.0e38f;
        s_conv[pos * 16 + j] = __float2bfloat16(y);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ST_PH * ST_PW * 16; i += ST_THREADS) {
      const int j = i & 15, pp = i >> 4;
      const int py = pp / ST_PW, px = pp % ST_PW;
      if (py0 + py < Hp && px0 + px < Wp) {
        float m = -3.0e38f;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int s = 0; s < 3; ++s) m = fmaxf(m, __bfloat162float(s_conv[((2 * py + r) * ST_CW + 2 * px + s) * 16 + j]));
        out[(((int64_t)b * Hp + py0 + py) * Wp + px0 + px) * 64 + cg * 16 + j] = __float2bfloat16(m);
      }
    }
    __syncthreads();
  }
}

// test tap "fuse" in the fused layout: rebuild cat[p5^8, p4^4, p3^2, p2] as NCHW fp32 from p2 and cat3 = [p5^4 | p4^2 | p3]
__global__ void fuse_tap_kernel(const __nv_bfloat16 *__restrict__ p2, const __nv_bfloat16 *__restrict__ cat3, int B, int H, int W,
                                float *__restrict__ out) {
  const int64_t n = (int64_t)B * 256 * H * W;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % W), y = (int)((i / W) % H), c = (int)((i / ((int64_t)W * H)) % 256), b = (int)(i / ((int64_t)W * H * 256));
  float v;
  if (c < 192) v = __bfloat162float(cat3[(((int64_t)b * (H / 2) + y / 2) * (W / 2) + x / 2) * 192 + c]);
  else v = __bfloat162float(p2[(((int64_t)b * H + y) * W + x) * 64 + (c - 192)]);
  out[i] = v;
}

__global__ void bf16_nhwc_to_nchw_f32_kernel(const __nv_bfloat16 *__restrict__ in, int B, int H, int W, int C, int ldc,
                                             float *__restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)B * C * H * W;
  if (idx >= total) return;
  int x = (int)(idx % W), y = (int)((idx / W) % H), c = (int)((idx / ((int64_t)W * H)) % C);
  int64_t b = idx / ((int64_t)W * H * C);
  out[idx] = __bfloat162float(in[((b * H + y) * W + x) * ldc + c]);
}

// ---------------------------------------------------------------------------------------
// host-side weight preparation
// ---------------------------------------------------------------------------------------
struct HostWeights {
  std::map<std::string, std::vector<float>> t;
  const std::vector<float> *get(const std::string &n) const {
    auto it = t.find(n);
    return it == t.end() ? nullptr : &it->second;
  }
};

struct ConvSpec {
  std::string name, bn;  // weight name prefix, bn prefix ("" = none)
  int cin, cout, k, stride, pad;
};

struct DevConv {
  int cin = 0, cout = 0, k = 1, stride = 1, pad = 0;
  bool has_bn = false;
  int ds_cin = 0;  // > 0: w16 carries a fused 1x1 stride-2 downsample of ds_cin channels after the 9 taps
  int tap_mask = 0x1ff;  // 3x3 taps with non-zero weights (class kernels of the fused FPN level use 4 of 9)
  bool pair = false;     // conv_halo pair mode: 128 rows = two parity-class kernels sharing one input patch
  DevBuf w32;    // fp32 [k*k][cin][cout]
  DevBuf w16;    // bf16 [cout][k*k*cin]
  DevBuf wsplit;  // FP32-accuracy mode on the tensor cores: bf16 term blocks [cout][tap][cin / 64][split_nblk][64] (conv_tc.cuh)
  int split_nblk = 0;
  DevBuf scale, shift;
  std::vector<float> scale_h, shift_h;  // host copies: kernel-parameter constants of the TMA-store epilogue
};

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}

template <class T>
static int upload(DevBuf &buf, const std::vector<T> &v) {
  OCRB_TRY(buf.reserve(v.size() * sizeof(T)));
  OCRB_CUDA(cudaMemcpy(buf.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return OCRB_OK;
}

static int fold_bn(const HostWeights &hw, const std::string &bn, int c, const std::vector<float> *conv_bias,
                   std::vector<float> &scale, std::vector<float> &shift) {
  scale.assign(c, 1.0f);
  shift.assign(c, 0.0f);
  if (bn.empty()) {
    if (conv_bias) shift = *conv_bias;
    return OCRB_OK;
  }
  const auto *g = hw.get(bn + ".weight"), *b = hw.get(bn + ".bias"), *m = hw.get(bn + ".running_mean"), *v = hw.get(bn + ".running_var");
  OCRB_REQUIRE(g && b && m && v, "missing batch-norm tensors for %s", bn.c_str());
  OCRB_REQUIRE((int)g->size() == c && (int)b->size() == c && (int)m->size() == c && (int)v->size() == c, "bad batch-norm size for %s", bn.c_str());
  for (int i = 0; i < c; ++i) {
    float s = (*g)[i] / sqrtf((*v)[i] + BN_EPS);
    scale[i] = s;
    float cb = conv_bias ? (*conv_bias)[i] : 0.0f;
    shift[i] = (*b)[i] + (cb - (*m)[i]) * s;
  }
  return OCRB_OK;
}

static int prep_conv_body(const HostWeights *hw, const std::vector<float> *w, const ConvSpec &sp, bool want16, DevConv &dc);
static int prep_conv(const HostWeights &hw, const ConvSpec &sp, bool want16, DevConv &dc) {
  const auto *w = hw.get(sp.name + ".weight");
  OCRB_REQUIRE(w, "missing weight tensor %s.weight", sp.name.c_str());
  return prep_conv_body(&hw, w, sp, want16, dc);
}
// the same from an OIHW weight vector that is not a VarStore tensor (no batch-norm lookup: identity affine)
static int prep_conv_from(const std::vector<float> &w, const ConvSpec &sp, bool want16, DevConv &dc) { return prep_conv_body(nullptr, &w, sp, want16, dc); }
static int prep_conv_body(const HostWeights *hwp, const std::vector<float> *w, const ConvSpec &sp, bool want16, DevConv &dc) {
  const int kk = sp.k * sp.k;
  OCRB_REQUIRE((int64_t)w->size() == (int64_t)sp.cout * sp.cin * kk, "tensor %s.weight has %zu elements, expected %lld",
               sp.name.c_str(), w->size(), (long long)sp.cout * sp.cin * kk);
  dc.cin = sp.cin; dc.cout = sp.cout; dc.k = sp.k; dc.stride = sp.stride; dc.pad = sp.pad;
  dc.has_bn = !sp.bn.empty();
  std::vector<float> scale(sp.cout, 1.0f), shift(sp.cout, 0.0f);
  if (hwp) OCRB_TRY(fold_bn(*hwp, sp.bn, sp.cout, nullptr, scale, shift));
  OCRB_TRY(upload(dc.scale, scale));
  OCRB_TRY(upload(dc.shift, shift));
  dc.scale_h = scale; dc.shift_h = shift;
  // OIHW -> [tap][ci][co] fp32
  std::vector<float> w32((size_t)kk * sp.cin * sp.cout);
  for (int co = 0; co < sp.cout; ++co)
    for (int ci = 0; ci < sp.cin; ++ci)
      for (int tp = 0; tp < kk; ++tp) w32[((size_t)tp * sp.cin + ci) * sp.cout + co] = (*w)[((size_t)co * sp.cin + ci) * kk + tp];
  OCRB_TRY(upload(dc.w32, w32));
  // FP32 mode: the operands of every 64-channel-aligned convolution as bf16 terms for the tensor cores
  // (OCRB_FP32=cuda keeps the CUDA-core kernels; OCRB_SPLIT_TERMS=3 selects the 24-bit form)
  static const bool fp32_cuda = getenv("OCRB_FP32") && strcmp(getenv("OCRB_FP32"), "cuda") == 0;
  static const int split_terms = getenv("OCRB_SPLIT_TERMS") && atoi(getenv("OCRB_SPLIT_TERMS")) == 3 ? 3 : 2;
  if (!want16 && !fp32_cuda && sp.cin % 64 == 0) {
    const int nblk = split_terms == 2 ? 3 : 6;
    static const int wplane2[3] = {0, 1, 0}, wplane3[6] = {0, 1, 2, 0, 1, 0};  // weight term of each K block (conv_tc.cuh)
    const int *wplane = split_terms == 2 ? wplane2 : wplane3;
    std::vector<uint16_t> ws((size_t)sp.cout * kk * sp.cin * nblk);
    for (int co = 0; co < sp.cout; ++co)
      for (int tp = 0; tp < kk; ++tp)
        for (int ci = 0; ci < sp.cin; ++ci) {
          float r = (*w)[((size_t)co * sp.cin + ci) * kk + tp];
          uint16_t term[3];
          for (int t = 0; t < 3; ++t) {
            term[t] = f2bf(r);
            uint32_t u = (uint32_t)term[t] << 16;
            float back;
            memcpy(&back, &u, 4);
            r -= back;
          }
          const size_t base = (((size_t)co * kk + tp) * (sp.cin / 64) + ci / 64) * nblk * 64 + ci % 64;
          for (int b = 0; b < nblk; ++b) ws[base + (size_t)b * 64] = term[wplane[b]];
        }
    OCRB_TRY(upload(dc.wsplit, ws));
    dc.split_nblk = nblk;
  }
  if (want16 && sp.cin % 64 == 0) {
    // OIHW -> [co][tap][ci] bf16 (K-major rows)
    std::vector<uint16_t> w16((size_t)sp.cout * kk * sp.cin);
    for (int co = 0; co < sp.cout; ++co)
      for (int tp = 0; tp < kk; ++tp)
        for (int ci = 0; ci < sp.cin; ++ci) w16[((size_t)co * kk + tp) * sp.cin + ci] = f2bf((*w)[((size_t)co * sp.cin + ci) * kk + tp]);
    OCRB_TRY(upload(dc.w16, w16));
  }
  return OCRB_OK;
}

// ResNet downsample (model.rs:30-38) folded into the block's conv2:
//   relu(bn2(conv2(t)) + bn_ds(conv_ds(x))) = relu(s2 * (W2*t + (s_ds/s2) W_ds*x) + t2 + t_ds)
// so the 1x1 stride-2 convolution becomes cin_ds extra K columns of conv2's weight matrix
// (pre-scaled per output channel in fp32, then rounded to bf16) and its shift joins conv2's.
static int prep_fused_downsample(const HostWeights &hw, const std::string &p, int cin_ds, int c, DevConv &dc) {
  const auto *w2 = hw.get(p + ".conv2.weight"), *wd = hw.get(p + ".downsample.0.weight");
  OCRB_REQUIRE(w2 && wd, "missing conv2 / downsample weights of %s", p.c_str());
  std::vector<float> s2, t2, sd, td;
  OCRB_TRY(fold_bn(hw, p + ".bn2", c, nullptr, s2, t2));
  OCRB_TRY(fold_bn(hw, p + ".downsample.1", c, nullptr, sd, td));
  for (int co = 0; co < c; ++co)
    if (!(fabsf(s2[co]) > 1e-20f)) return OCRB_OK;  // bn2 scale of 0: keep the separate downsample launch
  const int ktot = 9 * c + cin_ds;
  std::vector<uint16_t> w16((size_t)c * ktot);
  for (int co = 0; co < c; ++co) {
    for (int tp = 0; tp < 9; ++tp)
      for (int ci = 0; ci < c; ++ci) w16[(size_t)co * ktot + tp * c + ci] = f2bf((*w2)[((size_t)co * c + ci) * 9 + tp]);
    const float ratio = sd[co] / s2[co];
    for (int ci = 0; ci < cin_ds; ++ci) w16[(size_t)co * ktot + 9 * c + ci] = f2bf((*wd)[(size_t)co * cin_ds + ci] * ratio);
    t2[co] += td[co];
  }
  OCRB_TRY(upload(dc.w16, w16));
  OCRB_TRY(upload(dc.shift, t2));
  dc.shift_h = t2;  // the host copy feeds the TMA-store epilogue's kernel-parameter constants: keep it in step
  dc.scale_h = s2;
  dc.ds_cin = cin_ds;
  return OCRB_OK;
}

// FPN level 2 without its 256-channel intermediate (model.rs:126-129, BF16 mode):
//   p2 = out2(up2(in3) + in2(x1)),  in2 = 1x1 conv without bias, out2 = 3x3 conv without bias, zero padding
//      = conv3x3(W_out2 o W_in2)(x1)  +  conv3x3(W_out2)(up2(in3))
// * the first term is a 64 -> 64 3x3 convolution of x1 with the composed weights (K = 576 instead of 2304);
// * the second, a 3x3 convolution of a nearest-upsampled map, is for each output parity class (a, b) = (y & 1, x & 1)
//   a 2x2 convolution of the half-resolution lateral itself (and, in3 being a 1x1 convolution of x2, of x2 with weights
//   composed once more): row offsets {-1, 0} carry W[-1], W[0]+W[1] for even y and
//   {0, +1} carry W[-1]+W[0], W[1] for odd y (same for columns) — four 4-tap convolutions at 100 x 100 whose results are
//   stored pixel-shuffled (class (a,b) of low-res pixel (Y,X) at (2Y+a, 2X+b)) into one 64-channel map that the first
//   convolution adds as a residual.  Zero padding maps to zero padding, so borders are exact.
// 13.1 -> 8.2 GFLOP per image for this level and s2 (20 MB per image written and re-read) never exists.
static void compose_weights(const std::vector<float> &w_out /*[co][m][3][3]*/, const std::vector<float> &w_in /*[m][ci]*/, int co_n, int m_n,
                            int ci_n, std::vector<float> &wc /*[co][ci][3][3]*/) {
  wc.assign((size_t)co_n * ci_n * 9, 0.0f);
  std::vector<double> acc((size_t)ci_n);
  for (int co = 0; co < co_n; ++co)
    for (int tp = 0; tp < 9; ++tp) {
      std::fill(acc.begin(), acc.end(), 0.0);
      for (int m = 0; m < m_n; ++m) {
        const double wo = w_out[((size_t)co * m_n + m) * 9 + tp];
        const float *wi = &w_in[(size_t)m * ci_n];
        for (int ci = 0; ci < ci_n; ++ci) acc[ci] += wo * wi[ci];
      }
      for (int ci = 0; ci < ci_n; ++ci) wc[((size_t)co * ci_n + ci) * 9 + tp] = (float)acc[ci];
    }
}

// OIHW fp32 -> DevConv (bf16 K-major rows [co][tap][ci]) for a bias-free, BN-free 3x3 convolution
static int upload_plain_conv3(const std::vector<float> &w, int cin, int cout, int tap_mask, DevConv &dc) {
  dc.cin = cin; dc.cout = cout; dc.k = 3; dc.stride = 1; dc.pad = 1; dc.has_bn = false; dc.tap_mask = tap_mask;
  std::vector<uint16_t> w16((size_t)cout * 9 * cin);
  for (int co = 0; co < cout; ++co)
    for (int tp = 0; tp < 9; ++tp)
      for (int ci = 0; ci < cin; ++ci) w16[((size_t)co * 9 + tp) * cin + ci] = f2bf(w[((size_t)co * cin + ci) * 9 + tp]);
  return upload(dc.w16, w16);
}

// Classes (a, 1) at low-res column X and (a, 0) at column X + 1 read the SAME 2x2 input patch (columns X, X + 1): one
// 128-row weight matrix = [class (a,1) | class (a,0) shifted one tap to the right], taps rows(a) x {1, 2}, feeds both from
// one operand tile (N = 128: the tensor pipe is no longer starved by shared-memory operand reads as with N = 64).
// conv_halo's pair mode stores half 0 at column X and half 1 at column X + 1 of their pixel-shuffled positions.
static int upload_pair_conv3(const std::vector<float> &w_b1, const std::vector<float> &w_b0, int cin, int a, DevConv &dc) {
  std::vector<float> wp((size_t)128 * cin * 9, 0.0f);
  int mask = 0;
  for (int rr = a; rr <= a + 1; ++rr)  // rows(a): {0, 1} for a = 0, {1, 2} for a = 1
    for (int ss = 1; ss <= 2; ++ss) {
      mask |= 1 << (rr * 3 + ss);
      for (int co = 0; co < 64; ++co)
        for (int ci = 0; ci < cin; ++ci) {
          wp[((size_t)co * cin + ci) * 9 + rr * 3 + ss] = w_b1[((size_t)co * cin + ci) * 9 + rr * 3 + ss];
          wp[((size_t)(64 + co) * cin + ci) * 9 + rr * 3 + ss] = w_b0[((size_t)co * cin + ci) * 9 + rr * 3 + ss - 1];
        }
    }
  OCRB_TRY(upload_plain_conv3(wp, cin, 128, mask, dc));
  dc.pair = true;
  return OCRB_OK;
}

// one FPN level: out_l(in_l(x_l) + up2(in_u(x_u))) -> "<out>.x" (3x3 on x_l, composed) + "<out>.up{a}{b}" (4-tap classes on x_u)
static int prep_fused_fpn_level(const HostWeights &hw, const std::string &out, const std::string &in_l, const std::string &in_u, int c_l, int c_u,
                                std::map<std::string, DevConv> &conv) {
  const auto *wo = hw.get(out + ".weight"), *wi = hw.get(in_l + ".weight"), *wu = hw.get(in_u + ".weight");
  OCRB_REQUIRE(wo && wi && wu && wo->size() == (size_t)64 * 256 * 9 && wi->size() == (size_t)256 * c_l && wu->size() == (size_t)256 * c_u,
               "missing / mis-shaped %s / %s / %s weights", out.c_str(), in_l.c_str(), in_u.c_str());
  std::vector<float> wc;
  compose_weights(*wo, *wi, 64, 256, c_l, wc);
  OCRB_TRY(upload_plain_conv3(wc, c_l, 64, 0x1ff, conv[out + ".x"]));
  std::vector<float> cls[2][2];
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      // low-res tap r' (row offset r' - 1) collects the full-res taps dy whose source row (2Y + a + dy) >> 1 is Y + r' - 1
      std::vector<float> wk((size_t)64 * 256 * 9, 0.0f);
      int mask = 0;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int rr = ((a + dy) >> 1) + 1, ss = ((b + dx) >> 1) + 1;  // arithmetic shift = floor
          mask |= 1 << (rr * 3 + ss);
          for (int co = 0; co < 64; ++co)
            for (int m = 0; m < 256; ++m) wk[((size_t)co * 256 + m) * 9 + rr * 3 + ss] += (*wo)[((size_t)co * 256 + m) * 9 + (dy + 1) * 3 + dx + 1];
        }
      // the upper lateral is itself a bias-free 1x1 convolution of x_u: compose once more (K = 4 x c_u), and the raw
      // lateral is never materialised for this level
      std::vector<float> wkc;
      compose_weights(wk, *wu, 64, 256, c_u, wkc);
      OCRB_TRY(upload_plain_conv3(wkc, c_u, 64, mask, conv[out + ".up" + std::to_string(a) + std::to_string(b)]));
      cls[a][b].swap(wkc);
    }
  for (int a = 0; a < 2; ++a) OCRB_TRY(upload_pair_conv3(cls[a][1], cls[a][0], c_u, a, conv[out + ".pair" + std::to_string(a)]));
  return OCRB_OK;
}

static int prep_fused_fpn2(const HostWeights &hw, std::map<std::string, DevConv> &conv) {
  OCRB_TRY(prep_fused_fpn_level(hw, "out2", "in2", "in3", 64, 128, conv));   // p2 from x1 (200^2) and x2
  return prep_fused_fpn_level(hw, "out3", "in3", "in4", 128, 256, conv);     // p3 from x2 (100^2) and x3
}

// The same parity-class identity shrinks the concat buffer (model.rs:140-143).  Nearest upsampling composes
// (up8 = up2 o up4, up4 = up2 o up2), so cat[p5^8, p4^4, p3^2] = up2(cat3) with cat3 = cat[p5^4, p4^2, p3] at 100 x 100 and
//   bin_conv1(cat[p5^8, p4^4, p3^2, p2]) = conv3x3(W_bin[:, p2 channels])(p2)  +  conv3x3(W_bin[:, 0..191])(up2(cat3)),
// the second term again four 4-tap class convolutions at 100 x 100 (K = 4 x 192), stored pixel-shuffled and added as a
// residual: 768 instead of 1728 MACs per output pixel and channel, and the 256-channel 200 x 200 buffer is never built.
// bin_bn1 multiplies the whole sum, so its scale is folded into the class weights (the residual joins after the main
// convolution's scale / shift).
static int prep_fused_bin_p3(const HostWeights &hw, std::map<std::string, DevConv> &conv) {
  const auto *wb = hw.get("bin_conv1.weight");
  OCRB_REQUIRE(wb && wb->size() == (size_t)64 * 256 * 9, "missing / mis-shaped bin_conv1 weights");
  std::vector<float> sc, sh;
  OCRB_TRY(fold_bn(hw, "bin_bn1", 64, nullptr, sc, sh));
  // main part: p2 = reference channels 192..255
  std::vector<float> wm((size_t)64 * 64 * 9);
  for (int co = 0; co < 64; ++co)
    for (int ci = 0; ci < 64; ++ci)
      for (int tp = 0; tp < 9; ++tp) wm[((size_t)co * 64 + ci) * 9 + tp] = (*wb)[((size_t)co * 256 + 192 + ci) * 9 + tp];
  DevConv &m = conv["bin_conv1.main"];
  OCRB_TRY(upload_plain_conv3(wm, 64, 64, 0x1ff, m));
  m.has_bn = true;
  OCRB_TRY(upload(m.scale, sc));
  OCRB_TRY(upload(m.shift, sh));
  m.scale_h = sc; m.shift_h = sh;
  std::vector<float> cls[2][2];
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      std::vector<float> wk((size_t)64 * 192 * 9, 0.0f);
      int mask = 0;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int rr = ((a + dy) >> 1) + 1, ss = ((b + dx) >> 1) + 1;
          mask |= 1 << (rr * 3 + ss);
          for (int co = 0; co < 64; ++co)
            for (int ci = 0; ci < 192; ++ci)  // cat3 keeps the reference channel order [p5 | p4 | p3]
              wk[((size_t)co * 192 + ci) * 9 + rr * 3 + ss] += sc[co] * (*wb)[((size_t)co * 256 + ci) * 9 + (dy + 1) * 3 + dx + 1];
        }
      OCRB_TRY(upload_plain_conv3(wk, 192, 64, mask, conv["bin_conv1.up" + std::to_string(a) + std::to_string(b)]));
      cls[a][b].swap(wk);
    }
  for (int a = 0; a < 2; ++a) OCRB_TRY(upload_pair_conv3(cls[a][1], cls[a][0], 192, a, conv["bin_conv1.pair" + std::to_string(a)]));
  return OCRB_OK;
}

}  // namespace ocrb

using namespace ocrb;

struct ocrb_det {
  ocrb_ctx *ctx = nullptr;
  int mode = OCRB_MODE_BF16;
  // stem
  DevBuf stem_w, stem_scale, stem_shift;
  float stem_scale_h[64], stem_shift_h[64];  // host copies (kernel-parameter constants of stem_tc)
  // body: index by name
  std::map<std::string, DevConv> conv;
  // head
  DevBuf tr1_w32, tr1_scale, tr1_shift;  // fp32 path: [4][64][64]
  DevBuf head_w16;                       // bf16 path: [256][64]
  DevBuf tr2_w;                          // [4][64] fp32 (both paths)
  float tr2_bias = 0.f;
  HeadConsts head_c;                     // bf16 path: head constants passed as a kernel parameter
  // activations (grow-only), keyed by name
  std::map<std::string, DevBuf> act;
  DevBuf staged_in, staged_out;
  PinBuf err;  // pipeline-timeout code: pinned host memory written by the kernels, readable even after a trap killed the context
  // last forward (for taps)
  int last_B = 0, last_H = 0, last_W = 0;
  // BF16 mode: level 2 of the FPN computed from x1 and in3 directly (prep_fused_fpn2) and p3 kept out of the concat
  // buffer (prep_fused_bin_p3): "b.fuse" is then just p2 [B][H/4][W/4][64], "b.cat3" [B][H/8][W/8][192] = [p5^4 | p4^2 | p3]
  bool fpn2_fused = false;
  // tensor-map cache
  struct Maps { int B = 0, H = 0, W = 0; std::map<std::string, CUtensorMap> m; } maps;
};

namespace ocrb {

static const char *LAYER_NAMES[] = {"layer1", "layer2", "layer3", "layer4"};
static const int LAYER_C[] = {64, 128, 256, 512};

static int det_build(ocrb_det *d, const HostWeights &hw) {
  const bool bf = d->mode == OCRB_MODE_BF16;
  // stem
  {
    const auto *w = hw.get("conv1.weight");
    OCRB_REQUIRE(w && w->size() == 64 * 49, "missing or mis-sized conv1.weight");
    std::vector<float> wt(49 * 64), sc, sh;
    for (int co = 0; co < 64; ++co)
      for (int tp = 0; tp < 49; ++tp) wt[tp * 64 + co] = (*w)[co * 49 + tp];
    OCRB_TRY(fold_bn(hw, "bn1", 64, nullptr, sc, sh));
    memcpy(d->stem_scale_h, sc.data(), sizeof(d->stem_scale_h));
    memcpy(d->stem_shift_h, sh.data(), sizeof(d->stem_shift_h));
    OCRB_TRY(upload(d->stem_w, wt));
    OCRB_TRY(upload(d->stem_scale, sc));
    OCRB_TRY(upload(d->stem_shift, sh));
  }
  int cin = 64;
  for (int li = 0; li < 4; ++li) {
    const int c = LAYER_C[li];
    for (int blk = 0; blk < 2; ++blk) {
      std::string p = std::string(LAYER_NAMES[li]) + "." + std::to_string(blk);
      const int bc_in = blk == 0 ? cin : c;
      const int stride = (blk == 0 && li > 0) ? 2 : 1;
      OCRB_TRY(prep_conv(hw, {p + ".conv1", p + ".bn1", bc_in, c, 3, stride, 1}, bf, d->conv[p + ".conv1"]));
      OCRB_TRY(prep_conv(hw, {p + ".conv2", p + ".bn2", c, c, 3, 1, 1}, bf, d->conv[p + ".conv2"]));
      if (blk == 0 && li > 0) {
        OCRB_TRY(prep_conv(hw, {p + ".downsample.0", p + ".downsample.1", bc_in, c, 1, stride, 0}, bf, d->conv[p + ".downsample"]));
        static const bool fuse_ds = !(getenv("OCRB_FUSE_DS") && atoi(getenv("OCRB_FUSE_DS")) == 0);
        if (bf && fuse_ds) {
          // keep an unfused copy of conv2 for callers that cannot use the fused form (none today), fuse into conv2
          OCRB_TRY(prep_fused_downsample(hw, p, bc_in, c, d->conv[p + ".conv2"]));
        }
      }
    }
    cin = c;
  }
  OCRB_TRY(prep_conv(hw, {"in5", "", 512, 256, 1, 1, 0}, bf, d->conv["in5"]));
  OCRB_TRY(prep_conv(hw, {"in4", "", 256, 256, 1, 1, 0}, bf, d->conv["in4"]));
  OCRB_TRY(prep_conv(hw, {"in3", "", 128, 256, 1, 1, 0}, bf, d->conv["in3"]));
  OCRB_TRY(prep_conv(hw, {"in2", "", 64, 256, 1, 1, 0}, bf, d->conv["in2"]));
  for (const char *n : {"out5", "out4", "out3", "out2"}) OCRB_TRY(prep_conv(hw, {n, "", 256, 64, 3, 1, 1}, bf, d->conv[n]));
  OCRB_TRY(prep_conv(hw, {"bin_conv1", "bin_bn1", 256, 64, 3, 1, 1}, bf, d->conv["bin_conv1"]));
  {
    static const bool fuse_fpn2 = !(getenv("OCRB_FUSE_FPN2") && atoi(getenv("OCRB_FUSE_FPN2")) == 0);  // tuning / bisecting knob
    static const bool halo_on = !(getenv("OCRB_CONV") && strcmp(getenv("OCRB_CONV"), "tc") == 0);
    if (bf && fuse_fpn2 && halo_on && halo_use_ts(64, 1)) {
      OCRB_TRY(prep_fused_fpn2(hw, d->conv));
      OCRB_TRY(prep_fused_bin_p3(hw, d->conv));
      d->fpn2_fused = true;
    }
  }
  // head: conv-transpose weights are [in][out][kh][kw]
  {
    const auto *w1 = hw.get("bin_conv_tr1.weight"), *b1 = hw.get("bin_conv_tr1.bias");
    const auto *w2 = hw.get("bin_conv_tr2.weight"), *b2 = hw.get("bin_conv_tr2.bias");
    OCRB_REQUIRE(w1 && b1 && w2 && b2, "missing bin_conv_tr1/2 tensors");
    OCRB_REQUIRE(w1->size() == 64 * 64 * 4 && b1->size() == 64 && w2->size() == 64 * 4 && b2->size() == 1, "bad bin_conv_tr shapes");
    std::vector<float> sc, sh;
    OCRB_TRY(fold_bn(hw, "bin_bn2", 64, b1, sc, sh));
    OCRB_TRY(upload(d->tr1_scale, sc));
    OCRB_TRY(upload(d->tr1_shift, sh));
    std::vector<float> w32(4 * 64 * 64);  // [tap][ci][co]
    std::vector<uint16_t> w16(256 * 64);  // [tap*64+co][ci]
    for (int ci = 0; ci < 64; ++ci)
      for (int co = 0; co < 64; ++co)
        for (int tp = 0; tp < 4; ++tp) {
          float v = (*w1)[((size_t)ci * 64 + co) * 4 + tp];
          w32[((size_t)tp * 64 + ci) * 64 + co] = v;
          w16[((size_t)tp * 64 + co) * 64 + ci] = f2bf(v);
        }
    OCRB_TRY(upload(d->tr1_w32, w32));
    if (bf) OCRB_TRY(upload(d->head_w16, w16));
    if (!bf) {
      // FP32 mode on the tensor cores: conv-transpose 1 as a 1x1 convolution 64 -> 4 * 64 (column = tap * 64 + co) through the
      // split-operand engine; bin_bn2 (with the bias folded in) repeats per tap
      DevConv &c = d->conv["tr1"];
      std::vector<float> oihw((size_t)256 * 64), sc4(256), sh4(256);
      for (int tp = 0; tp < 4; ++tp)
        for (int co = 0; co < 64; ++co) {
          sc4[tp * 64 + co] = sc[co];
          sh4[tp * 64 + co] = sh[co];
          for (int ci = 0; ci < 64; ++ci) oihw[((size_t)(tp * 64 + co)) * 64 + ci] = (*w1)[((size_t)ci * 64 + co) * 4 + tp];
        }
      OCRB_TRY(prep_conv_from(oihw, {"tr1", "", 64, 256, 1, 1, 0}, false, c));
      OCRB_TRY(upload(c.scale, sc4));
      OCRB_TRY(upload(c.shift, sh4));
    }
    std::vector<float> w2t(4 * 64);  // [tap][ci]
    for (int ci = 0; ci < 64; ++ci)
      for (int tp = 0; tp < 4; ++tp) w2t[tp * 64 + ci] = (*w2)[ci * 4 + tp];
    OCRB_TRY(upload(d->tr2_w, w2t));
    d->tr2_bias = (*b2)[0];
    for (int i = 0; i < 64; ++i) { d->head_c.scale[i] = sc[i]; d->head_c.shift[i] = sh[i]; }
    for (int q = 0; q < 4; ++q)
      for (int co = 0; co < 64; ++co) d->head_c.w2[co * 4 + q] = w2t[q * 64 + co];
  }
  OCRB_TRY(d->err.reserve(4));
  *d->err.as<int>() = 0;
  return OCRB_OK;
}

template <class T>
static int act(ocrb_det *d, const std::string &name, int64_t elems, T **out) {
  DevBuf &b = d->act[name];
  OCRB_TRY(b.reserve((size_t)elems * sizeof(T)));
  *out = b.as<T>();
  return OCRB_OK;
}

static int n_tile_for(int cout);

// ------------------------------------------------------------------------------ FP32 graph
static int forward_fp32(ocrb_det *d, const float *img /*[B][H][W] dev*/, int B, int H, int W, float *prob) {
  ocrb_ctx *ctx = d->ctx;
  const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4;
  float *c1, *x0;
  OCRB_TRY(act(d, "f.c1", (int64_t)B * H2 * W2 * 64, &c1));
  OCRB_TRY(act(d, "f.stem", (int64_t)B * H4 * W4 * 64, &x0));
  OCRB_TRY(launch_stem_fp32(ctx, img, B, H, W, d->stem_w.as<float>(), d->stem_scale.as<float>(), d->stem_shift.as<float>(), c1));
  OCRB_TRY(launch_maxpool_fp32(ctx, c1, B, H2, W2, 64, x0));
  auto conv = [&](const std::string &name, const float *in, int h, int w, const float *res, int relu, float *out) -> int {
    DevConv &c = d->conv[name];
    if (c.split_nblk == 0)
      return launch_conv_fp32(ctx, in, B, h, w, c.cin, c.w32.as<float>(), c.cout, c.k, c.stride, c.pad, c.scale.as<float>(),
                              c.shift.as<float>(), res, relu, out);
    // tensor-core path at fp32-class accuracy: the input is split into bf16 terms (channel planes), the convolution is a
    // tcgen05 implicit GEMM over split_nblk K blocks per (tap, chunk), accumulation / affine / residual / output stay fp32
    const int terms = c.split_nblk == 3 ? 2 : 3;
    DevBuf &sb = d->act["f.split"];
    OCRB_TRY(sb.reserve((size_t)B * h * w * c.cin * terms * 2));
    OCRB_TRY(launch_split_terms(ctx, in, (int64_t)B * h * w, c.cin, terms, sb.p));
    CUtensorMap tmA, tmB;
    const int nt = n_tile_for(c.cout);
    OCRB_TRY(make_act_tensor_map(&tmA, sb.p, B, h, w, terms * c.cin, c.stride));
    OCRB_TRY(make_weight_tensor_map(&tmB, c.wsplit.p, c.cout, c.k * c.k * c.cin * c.split_nblk, nt));
    ConvTcParams p;
    p.B = B;
    p.Ho = (h + 2 * c.pad - c.k) / c.stride + 1;
    p.Wo = (w + 2 * c.pad - c.k) / c.stride + 1;
    p.Cout = c.cout;
    p.R = c.k; p.S = c.k; p.cin_chunks = c.split_nblk * (c.cin / 64); p.stride = c.stride; p.pad = c.pad;
    p.scale = c.scale.as<float>(); p.shift = c.shift.as<float>();
    p.res32 = res; p.relu = relu; p.out32 = out;
    p.split_nblk = c.split_nblk; p.split_cin = c.cin;
    p.err = d->err.as<int>();
    return launch_conv_tc(ctx, tmA, tmB, p, nt, EPI_F32, ("tc:" + name).c_str());
  };
  const float *x = x0;
  int h = H4, w = W4;
  const float *feat[4];
  int fh[4], fw[4];
  for (int li = 0; li < 4; ++li) {
    const int c = LAYER_C[li];
    const int ho = li > 0 ? h / 2 : h, wo = li > 0 ? w / 2 : w;
    for (int blk = 0; blk < 2; ++blk) {
      std::string p = std::string(LAYER_NAMES[li]) + "." + std::to_string(blk);
      float *t, *y, *ds = nullptr;
      OCRB_TRY(act(d, "f." + p + ".t", (int64_t)B * ho * wo * c, &t));
      OCRB_TRY(act(d, "f." + p + ".y", (int64_t)B * ho * wo * c, &y));
      const int hin = blk == 0 ? h : ho, win = blk == 0 ? w : wo;
      OCRB_TRY(conv(p + ".conv1", x, hin, win, nullptr, 1, t));
      const float *res = x;
      if (blk == 0 && li > 0) {
        OCRB_TRY(act(d, "f." + p + ".ds", (int64_t)B * ho * wo * c, &ds));
        OCRB_TRY(conv(p + ".downsample", x, hin, win, nullptr, 0, ds));
        res = ds;
      }
      OCRB_TRY(conv(p + ".conv2", t, ho, wo, res, 1, y));
      x = y;
    }
    h = ho; w = wo;
    feat[li] = x; fh[li] = h; fw[li] = w;
  }
  float *in5, *in4, *in3, *in2, *s4, *s3, *s2, *p5, *p4, *p3, *fuse, *b1, *t1;
  OCRB_TRY(act(d, "f.in5", (int64_t)B * fh[3] * fw[3] * 256, &in5));
  OCRB_TRY(act(d, "f.in4", (int64_t)B * fh[2] * fw[2] * 256, &in4));
  OCRB_TRY(act(d, "f.in3", (int64_t)B * fh[1] * fw[1] * 256, &in3));
  OCRB_TRY(act(d, "f.in2", (int64_t)B * fh[0] * fw[0] * 256, &in2));
  OCRB_TRY(act(d, "f.s4", (int64_t)B * fh[2] * fw[2] * 256, &s4));
  OCRB_TRY(act(d, "f.s3", (int64_t)B * fh[1] * fw[1] * 256, &s3));
  OCRB_TRY(act(d, "f.s2", (int64_t)B * fh[0] * fw[0] * 256, &s2));
  OCRB_TRY(act(d, "f.p5", (int64_t)B * fh[3] * fw[3] * 64, &p5));
  OCRB_TRY(act(d, "f.p4", (int64_t)B * fh[2] * fw[2] * 64, &p4));
  OCRB_TRY(act(d, "f.p3", (int64_t)B * fh[1] * fw[1] * 64, &p3));
  OCRB_TRY(act(d, "f.fuse", (int64_t)B * H4 * W4 * 256, &fuse));
  OCRB_TRY(act(d, "f.bin1", (int64_t)B * H4 * W4 * 64, &b1));
  OCRB_TRY(act(d, "f.t1", (int64_t)B * H2 * W2 * 64, &t1));
  OCRB_TRY(conv("in5", feat[3], fh[3], fw[3], nullptr, 0, in5));
  OCRB_TRY(conv("in4", feat[2], fh[2], fw[2], nullptr, 0, in4));
  OCRB_TRY(conv("in3", feat[1], fh[1], fw[1], nullptr, 0, in3));
  OCRB_TRY(conv("in2", feat[0], fh[0], fw[0], nullptr, 0, in2));
  OCRB_TRY(launch_upsample2_add_fp32(ctx, in5, in4, B, fh[2], fw[2], 256, s4));
  OCRB_TRY(launch_upsample2_add_fp32(ctx, in4, in3, B, fh[1], fw[1], 256, s3));
  OCRB_TRY(launch_upsample2_add_fp32(ctx, in3, in2, B, fh[0], fw[0], 256, s2));
  OCRB_TRY(conv("out5", in5, fh[3], fw[3], nullptr, 0, p5));
  OCRB_TRY(conv("out4", s4, fh[2], fw[2], nullptr, 0, p4));
  OCRB_TRY(conv("out3", s3, fh[1], fw[1], nullptr, 0, p3));
  // out2 writes straight into its concat slice?  conv_fp32 has no pitch: use a temp + copy
  float *p2;
  OCRB_TRY(act(d, "f.p2", (int64_t)B * H4 * W4 * 64, &p2));
  OCRB_TRY(conv("out2", s2, fh[0], fw[0], nullptr, 0, p2));
  OCRB_TRY(launch_upsample_concat_fp32(ctx, p5, B, H4, W4, 64, 8, 0, 256, fuse));
  OCRB_TRY(launch_upsample_concat_fp32(ctx, p4, B, H4, W4, 64, 4, 64, 256, fuse));
  OCRB_TRY(launch_upsample_concat_fp32(ctx, p3, B, H4, W4, 64, 2, 128, 256, fuse));
  OCRB_TRY(launch_upsample_concat_fp32(ctx, p2, B, H4, W4, 64, 1, 192, 256, fuse));
  OCRB_TRY(conv("bin_conv1", fuse, H4, W4, nullptr, 1, b1));
  if (d->conv.count("tr1") && d->conv["tr1"].split_nblk) {
    // t1 in tap-major form [B][H4][W4][4 taps][64]: pixel (2y + i, 2x + j) of the 400 x 400 map is tap 2i + j of pixel (y, x)
    OCRB_TRY(conv("tr1", b1, H4, W4, nullptr, 1, t1));
    OCRB_TRY(launch_convt2x2_sigmoid_fp32(ctx, t1, B, H2, W2, 64, d->tr2_w.as<float>(), d->tr2_bias, prob, 1));
  } else {
    OCRB_TRY(launch_convt2x2_fp32(ctx, b1, B, H4, W4, 64, 64, d->tr1_w32.as<float>(), d->tr1_scale.as<float>(),
                                  d->tr1_shift.as<float>(), 1, t1));
    OCRB_TRY(launch_convt2x2_sigmoid_fp32(ctx, t1, B, H2, W2, 64, d->tr2_w.as<float>(), d->tr2_bias, prob, 0));
  }
  return OCRB_OK;
}

// ------------------------------------------------------------------------------ BF16 graph
static int get_map(ocrb_det *d, const std::string &key, const void *base, int B, int H, int W, int C, int stride, CUtensorMap **out) {
  auto it = d->maps.m.find(key);
  if (it == d->maps.m.end()) {
    CUtensorMap m;
    OCRB_TRY(make_act_tensor_map(&m, base, B, H, W, C, stride));
    it = d->maps.m.emplace(key, m).first;
  }
  *out = &it->second;
  return OCRB_OK;
}
static int get_wmap(ocrb_det *d, const std::string &key, const void *base, int Cout, int Ktot, int n_tile, CUtensorMap **out) {
  auto it = d->maps.m.find(key);
  if (it == d->maps.m.end()) {
    CUtensorMap m;
    OCRB_TRY(make_weight_tensor_map(&m, base, Cout, Ktot, n_tile));
    it = d->maps.m.emplace(key, m).first;
  }
  *out = &it->second;
  return OCRB_OK;
}

static int n_tile_for(int cout) { return cout >= 256 ? 256 : cout; }

template <class TIn>
static int forward_bf16(ocrb_det *d, const TIn *img, int B, int H, int W, float *prob, uint8_t *bitmap, float thresh) {
  typedef __nv_bfloat16 bf;
  ocrb_ctx *ctx = d->ctx;
  const int H4 = H / 4, W4 = W / 4;
  // All activation buffers are allocated first: cached tensor maps embed their addresses.
  int fh[4], fw[4];
  fh[0] = H4; fw[0] = W4;
  for (int i = 1; i < 4; ++i) { fh[i] = fh[i - 1] / 2; fw[i] = fw[i - 1] / 2; }
  size_t before = 0, after = 0;
  for (auto &kv : d->act) before += kv.second.cap;
  bf *x0;
  OCRB_TRY(act(d, "b.stem", (int64_t)B * H4 * W4 * 64, &x0));
  struct Blk { bf *t, *y, *ds; bool has_ds; };
  Blk blks[4][2];
  for (int li = 0; li < 4; ++li)
    for (int blk = 0; blk < 2; ++blk) {
      std::string p = std::string(LAYER_NAMES[li]) + "." + std::to_string(blk);
      const int64_t n = (int64_t)B * fh[li] * fw[li] * LAYER_C[li];
      OCRB_TRY(act(d, "b." + p + ".t", n, &blks[li][blk].t));
      OCRB_TRY(act(d, "b." + p + ".y", n, &blks[li][blk].y));
      blks[li][blk].ds = nullptr;
      blks[li][blk].has_ds = blk == 0 && li > 0;
      // the downsample output exists only when the 1x1 stride-2 convolution is NOT folded into conv2
      if (blks[li][blk].has_ds && d->conv[p + ".conv2"].ds_cin == 0) OCRB_TRY(act(d, "b." + p + ".ds", n, &blks[li][blk].ds));
    }
  bf *in5, *in4 = nullptr, *in3 = nullptr, *s4, *s3 = nullptr, *s2 = nullptr, *fuse, *b1;
  OCRB_TRY(act(d, "b.in5", (int64_t)B * fh[3] * fw[3] * 256, &in5));
  OCRB_TRY(act(d, "b.s4", (int64_t)B * fh[2] * fw[2] * 256, &s4));
  if (!d->fpn2_fused) {  // the fused neck reads x2 / x3 directly: raw in4, in3 and the sums s3 / s2 are never built
    OCRB_TRY(act(d, "b.in4", (int64_t)B * fh[2] * fw[2] * 256, &in4));
    OCRB_TRY(act(d, "b.in3", (int64_t)B * fh[1] * fw[1] * 256, &in3));
    OCRB_TRY(act(d, "b.s3", (int64_t)B * fh[1] * fw[1] * 256, &s3));
    OCRB_TRY(act(d, "b.s2", (int64_t)B * fh[0] * fw[0] * 256, &s2));
  }
  const int fuse_c = d->fpn2_fused ? 64 : 256;
  OCRB_TRY(act(d, "b.fuse", (int64_t)B * H4 * W4 * fuse_c, &fuse));
  OCRB_TRY(act(d, "b.bin1", (int64_t)B * H4 * W4 * 64, &b1));
  bf *up2 = nullptr;  // fused FPN level 2: out2's share of up2(in3), pixel-shuffled, [B][H4][W4][64]
  bf *p3 = nullptr;
  if (d->fpn2_fused) {
    OCRB_TRY(act(d, "b.up2", (int64_t)B * H4 * (W4 + 2) * 64, &up2));  // rows padded by one pixel at either end (class_convs)
    OCRB_TRY(act(d, "b.cat3", (int64_t)B * fh[1] * fw[1] * 192, &p3));
  }
  for (auto &kv : d->act) after += kv.second.cap;
  if (after != before || d->maps.B != B || d->maps.H != H || d->maps.W != W) {
    d->maps.m.clear();  // some buffer moved or the shape changed: rebuild the descriptors
    d->maps.B = B; d->maps.H = H; d->maps.W = W;
  }

  // stem
  static const bool stem_cuda_cores = getenv("OCRB_STEM") && strcmp(getenv("OCRB_STEM"), "cuda") == 0;
  // default: stem_tc3.cu (transposed implicit GEMM, pooling in registers).  OCRB_STEM=v1: the im2col stem of stem_tc.cu
  // (round 1: 8.2 ms per 1024 images against 4.1); v2: stem_tc2.cu (pixels in the TMEM lanes: issue-bound, 7.8 - 8.4 ms).
  // The u8 form of v3 fetches the image in aligned 16-byte chunks; other shapes take v1.
  static const bool stem_v1 = getenv("OCRB_STEM") && strcmp(getenv("OCRB_STEM"), "v1") == 0;
  static const bool stem_v2 = getenv("OCRB_STEM") && strcmp(getenv("OCRB_STEM"), "v2") == 0;
  const bool stem_v3 = !stem_v1 && !stem_v2 && (sizeof(TIn) != 1 || (W % 16 == 0 && (reinterpret_cast<uintptr_t>(img) & 15) == 0));
  if (!stem_cuda_cores && stem_v3) {
    OCRB_TRY(launch_stem_tc3(ctx, img, sizeof(TIn) == 1, B, H, W, d->stem_w.as<float>(), d->stem_scale_h, d->stem_shift_h, x0,
                             d->err.as<int>()));
  } else if (!stem_cuda_cores && stem_v2) {
    OCRB_TRY(launch_stem_tc2(ctx, img, sizeof(TIn) == 1, B, H, W, d->stem_w.as<float>(), d->stem_scale_h, d->stem_shift_h, x0,
                             d->err.as<int>()));
  } else if (!stem_cuda_cores) {
    OCRB_TRY(launch_stem_tc(ctx, img, sizeof(TIn) == 1, B, H, W, d->stem_w.as<float>(), d->stem_scale_h, d->stem_shift_h, x0,
                            d->err.as<int>()));
  } else {
    const int tiles = (int)(cdiv(W4, ST_PW) * cdiv(H4, ST_PH)) * B;
    stem_fused_bf16_kernel<TIn><<<tiles, ST_THREADS, 0, ctx->stream>>>(img, B, H, W, d->stem_w.as<float>(), d->stem_scale.as<float>(),
                                                                       d->stem_shift.as<float>(), x0);
    OCRB_TRY(check_launch(ctx, "stem_fused_bf16"));
  }
  static const bool use_halo = !(getenv("OCRB_CONV") && strcmp(getenv("OCRB_CONV"), "tc") == 0);
  auto conv = [&](const std::string &name, const bf *in, int h, int w, ConvTcParams p) -> int {
    DevConv &c = d->conv[name];
    CUtensorMap *ma, *mb;
    if (c.k == 1 && c.stride == 1 && p.sum_out && p.up_src && !c.has_bn && !p.relu && lateral_ts_supported(c.cin, c.cout, h, w))
      return launch_conv_lateral(ctx, in, c.w16.as<bf>(), p.up_src, p.out, p.sum_out, B, h, w, c.cin, d->err.as<int>(), ("tc:" + name).c_str());
    const bool halo = use_halo && c.k == 3 && c.stride == 1 && !p.sum_out && p.out;
    const int nt = halo ? (c.cout == 64 ? 64 : 128) : n_tile_for(c.cout);
    if (c.pair) {
      OCRB_REQUIRE(halo && p.out2, "pair convolution needs the halo kernel and two outputs");
      p.pair_mode = 1; p.in_x_off = -1;
    }
    if (halo) {
      auto it = d->maps.m.find("h." + name);
      if (it == d->maps.m.end()) {
        CUtensorMap m;
        OCRB_TRY(make_halo_act_map(&m, in, B, h, w, c.cin, nt, nt == 64 ? 4 : 2, p.rep, c.pair ? 1 : 0));
        it = d->maps.m.emplace("h." + name, m).first;
      }
      ma = &it->second;
    } else {
      OCRB_TRY(get_map(d, "a." + name, in, B, h, w, c.cin, c.stride, &ma));
    }
    OCRB_TRY(get_wmap(d, (halo ? "wh." : "w.") + name, c.w16.p, c.cout, c.k * c.k * c.cin + c.ds_cin, halo ? halo_weight_box_rows(nt) : nt, &mb));
    CUtensorMap *md = nullptr;
    if (c.ds_cin > 0) {  // fused downsample: second input = the block input sampled at stride 2
      OCRB_REQUIRE(halo && p.ds_src, "fused downsample needs the halo kernel and a source");
      auto it = d->maps.m.find("d." + name);
      if (it == d->maps.m.end()) {
        CUtensorMap m;
        OCRB_TRY(make_halo_ds_map(&m, p.ds_src, B, 2 * h, 2 * w, c.ds_cin, h, w, nt, nt == 64 ? 4 : 2));
        it = d->maps.m.emplace("d." + name, m).first;
      }
      md = &it->second;
      p.ds_chunks = c.ds_cin / 64;
    }
    p.B = B;
    p.Ho = (h + 2 * c.pad - c.k) / c.stride + 1;
    p.Wo = (w + 2 * c.pad - c.k) / c.stride + 1 + (c.pair ? 1 : 0);  // pair mode: output columns -1 .. w - 1
    p.Cout = c.cout;
    p.R = c.k; p.S = c.k; p.cin_chunks = c.cin / 64; p.stride = c.stride; p.pad = c.pad;
    p.scale = c.has_bn ? c.scale.as<float>() : nullptr;  // no batch-norm: identity epilogue
    p.shift = c.has_bn ? c.shift.as<float>() : nullptr;
    p.tap_mask = c.tap_mask;
    p.scale_host = c.has_bn && !c.scale_h.empty() ? c.scale_h.data() : nullptr;
    p.shift_host = c.has_bn && !c.shift_h.empty() ? c.shift_h.data() : nullptr;
    if (p.out && p.out_ldc == 0) p.out_ldc = c.cout;
    p.err = d->err.as<int>();
    const std::string tag = "tc:" + name;
    if (halo) return launch_conv_halo(ctx, *ma, *mb, p, nt, tag.c_str(), md);
    return launch_conv_tc(ctx, *ma, *mb, p, nt, EPI_STD, tag.c_str());
  };
  // the four parity-class convolutions of `in` (hl x wl), pixel-shuffled into dst [B][2 hl][2 wl][64]: two pair launches
  // (N = 128, classes (a,1) | (a,0) from one operand tile), or four single-class launches with OCRB_PAIR=0
  static const bool use_pairs = !(getenv("OCRB_PAIR") && atoi(getenv("OCRB_PAIR")) == 0);
  auto class_convs = [&](const std::string &prefix, const bf *in, int hl, int wl, bf *dst) -> int {
    // dst rows are padded: pixel x of a row sits at column x + 1 of 2 wl + 2 (TMA stores take no negative coordinates, and
    // a pair launch produces one surplus column at either end)
    const int64_t wp = 2 * (int64_t)wl + 2;
    for (int a = 0; a < 2; ++a) {
      if (use_pairs) {
        // output index i = low-res column + 1: class (a,1) of column i - 1 -> pixel 2i - 1 -> padded column 2i;
        // class (a,0) of column i -> pixel 2i -> padded column 2i + 1
        ConvTcParams k;
        k.out = dst + (a * wp) * 64; k.out2 = dst + (a * wp + 1) * 64; k.out_ldc = 64; k.out_step = 2; k.out_row_px = (int)wp;
        OCRB_TRY(conv(prefix + ".pair" + std::to_string(a), in, hl, wl, k));
      } else {
        for (int b = 0; b < 2; ++b) {
          ConvTcParams k;
          k.out = dst + (a * wp + b + 1) * 64; k.out_ldc = 64; k.out_step = 2; k.out_row_px = (int)wp;
          OCRB_TRY(conv(prefix + ".up" + std::to_string(a) + std::to_string(b), in, hl, wl, k));
        }
      }
    }
    return OCRB_OK;
  };
  const bf *x = x0;
  int h = H4, w = W4;
  const bf *feat[4];
  for (int li = 0; li < 4; ++li) {
    for (int blk = 0; blk < 2; ++blk) {
      std::string p = std::string(LAYER_NAMES[li]) + "." + std::to_string(blk);
      Blk &bk = blks[li][blk];
      const int hin = blk == 0 ? h : fh[li], win = blk == 0 ? w : fw[li];
      ConvTcParams q1;
      q1.relu = 1; q1.out = bk.t;
      OCRB_TRY(conv(p + ".conv1", x, hin, win, q1));
      const bf *res = x;
      const bool ds_fused = bk.has_ds && d->conv[p + ".conv2"].ds_cin > 0;
      if (bk.has_ds && !ds_fused) {
        ConvTcParams qd;
        qd.relu = 0; qd.out = bk.ds;
        OCRB_TRY(conv(p + ".downsample", x, hin, win, qd));
        res = bk.ds;
      }
      ConvTcParams q2;
      q2.relu = 1; q2.out = bk.y; q2.residual = ds_fused ? nullptr : res;
      q2.ds_src = ds_fused ? x : nullptr;
      OCRB_TRY(conv(p + ".conv2", bk.t, fh[li], fw[li], q2));
      x = bk.y;
    }
    h = fh[li]; w = fw[li];
    feat[li] = x;
  }
  {  // laterals with the FPN "+ up2" fused (the RAW lateral of the level above is added)
    ConvTcParams q;
    q.out = in5;
    OCRB_TRY(conv("in5", feat[3], fh[3], fw[3], q));
    q = ConvTcParams(); q.out = d->fpn2_fused ? nullptr : in4; q.up_src = in5; q.sum_out = s4;  // raw in4 only feeds the unfused in3
    OCRB_TRY(conv("in4", feat[2], fh[2], fw[2], q));
    if (!d->fpn2_fused) {  // fused levels 2 and 3 read x2 / x3 directly: neither in3 nor s3 exists
      q = ConvTcParams(); q.out = in3; q.up_src = in4; q.sum_out = s3;
      OCRB_TRY(conv("in3", feat[1], fh[1], fw[1], q));
    }
    if (!d->fpn2_fused) {
      q = ConvTcParams(); q.out = nullptr; q.up_src = in3; q.sum_out = s2;
      OCRB_TRY(conv("in2", feat[0], fh[0], fw[0], q));
    }
  }
  {  // out convs write their (replicated) result into the concat buffer: cat([p5,p4,p3,p2], 1)
    ConvTcParams q;
    // cat[p5, p4, p3] at 200 x 200 (x8, x4, x2), or — fused — at 100 x 100 (x4, x2, x1): bin_conv1 takes that one through its
    // class convolutions
    const int rdiv = d->fpn2_fused ? 2 : 1;
    q.out = d->fpn2_fused ? p3 : fuse; q.out_ldc = d->fpn2_fused ? 192 : 256;
    q.out_coff = 0; q.rep = 8 / rdiv;
    OCRB_TRY(conv("out5", in5, fh[3], fw[3], q));
    q.out_coff = 64; q.rep = 4 / rdiv;
    OCRB_TRY(conv("out4", s4, fh[2], fw[2], q));
    q.out_coff = 128; q.rep = 2 / rdiv;
    if (!d->fpn2_fused) {
      OCRB_TRY(conv("out3", s3, fh[1], fw[1], q));
    } else {
      // p3 = conv3x3(W_out3 o W_in3)(x2) + [class convolutions of x3, pixel-shuffled into the first quarter of up2]
      OCRB_TRY(class_convs("out3", feat[2], fh[2], fw[2], up2));
      q.residual = up2 + 64; q.res_row_px = fw[1] + 2;
      OCRB_TRY(conv("out3.x", feat[1], fh[1], fw[1], q));
      q.residual = nullptr; q.res_row_px = 0;
    }
    q.out = fuse; q.out_ldc = fuse_c;
    q.out_coff = d->fpn2_fused ? 0 : 192; q.rep = 1;
    if (!d->fpn2_fused) {
      OCRB_TRY(conv("out2", s2, fh[0], fw[0], q));
    } else {
      // p2 = conv3x3(W_out2 o W_in2)(x1) + [four 2x2 class convolutions of in3, pixel-shuffled into up2] (see prep_fused_fpn2)
      OCRB_TRY(class_convs("out2", feat[1], fh[1], fw[1], up2));
      q.residual = up2 + 64; q.res_row_px = fw[0] + 2;
      OCRB_TRY(conv("out2.x", feat[0], fh[0], fw[0], q));
    }
  }
  if (!d->fpn2_fused) {
    ConvTcParams q;
    q.relu = 1; q.out = b1;
    OCRB_TRY(conv("bin_conv1", fuse, H4, W4, q));
  } else {
    // up2 is free again (out2.x1 consumed it, same stream): it now collects the share of cat3 = [p5 | p4 | p3] in bin_conv1
    OCRB_TRY(class_convs("bin_conv1", p3, fh[1], fw[1], up2));
    ConvTcParams q;
    q.relu = 1; q.out = b1; q.residual = up2 + 64; q.res_row_px = fw[0] + 2;
    OCRB_TRY(conv("bin_conv1.main", fuse, H4, W4, q));
  }
  {  // head tail
    CUtensorMap *ma, *mb;
    OCRB_TRY(get_map(d, "a.head", b1, B, H4, W4, 64, 1, &ma));
    OCRB_TRY(get_wmap(d, "w.head", d->head_w16.p, 256, 64, 256, &mb));
    ConvTcParams q;
    q.B = B; q.Ho = H4; q.Wo = W4; q.Cout = 256; q.R = 1; q.S = 1; q.cin_chunks = 1; q.stride = 1; q.pad = 0;
    q.scale = d->tr1_scale.as<float>(); q.shift = d->tr1_shift.as<float>();
    q.w2 = d->tr2_w.as<float>(); q.b2 = d->tr2_bias; q.thresh = thresh;
    q.prob = prob; q.bitmap = bitmap; q.err = d->err.as<int>();
    OCRB_TRY(launch_conv_tc(ctx, *ma, *mb, q, 256, EPI_HEAD, "tc:head", &d->head_c));
  }
  return OCRB_OK;
}

int det_forward_device(ocrb_det *det, const void *img_dev, int dtype, int B, int H, int W, float *prob_dev, uint8_t *bitmap_dev,
                       float thresh) {
  ocrb_ctx *ctx = det->ctx;
  det->last_B = B; det->last_H = H; det->last_W = W;
  if (det->mode == OCRB_MODE_BF16) {
    if (dtype == OCRB_U8) return forward_bf16<uint8_t>(det, (const uint8_t *)img_dev, B, H, W, prob_dev, bitmap_dev, thresh);
    return forward_bf16<float>(det, (const float *)img_dev, B, H, W, prob_dev, bitmap_dev, thresh);
  }
  const float *fimg = (const float *)img_dev;
  if (dtype == OCRB_U8) {
    float *tmp;
    OCRB_TRY(act(det, "f.img", (int64_t)B * H * W, &tmp));
    OCRB_TRY(launch_u8_to_f32(ctx, (const uint8_t *)img_dev, (int64_t)B * H * W, 1.0f, tmp));
    fimg = tmp;
  }
  return forward_fp32(det, fimg, B, H, W, prob_dev);
}

ocrb_ctx *det_ctx(ocrb_det *det) { return det->ctx; }
int det_mode(ocrb_det *det) { return det->mode; }
int det_check_err(ocrb_det *det) {
  const int err = *reinterpret_cast<volatile int *>(det->err.p);
  if (err) { set_error("tcgen05 pipeline timeout (wait site %d)", err); return OCRB_ERR_INTERNAL; }
  return OCRB_OK;
}

}  // namespace ocrb

extern "C" {

int ocrb_det_create(ocrb_ctx *ctx, int n, const char *const *names, const float *const *data, const int64_t *numel, int mode,
                    ocrb_det **out) {
  OCRB_REQUIRE(ctx && names && data && numel && out && n > 0, "bad argument");
  OCRB_REQUIRE(mode == OCRB_MODE_FP32 || mode == OCRB_MODE_BF16, "unknown mode %d", mode);
  OCRB_CUDA(cudaSetDevice(ctx->device));
  HostWeights hw;
  for (int i = 0; i < n; ++i) {
    OCRB_REQUIRE(names[i] && data[i] && numel[i] >= 0, "bad tensor %d", i);
    hw.t[names[i]] = std::vector<float>(data[i], data[i] + numel[i]);
  }
  ocrb_det *d = new ocrb_det();
  d->ctx = ctx;
  d->mode = mode;
  int rc = det_build(d, hw);
  if (rc != OCRB_OK) {
    ocrb_det_destroy(d);
    return rc;
  }
  *out = d;
  return OCRB_OK;
}

int ocrb_det_destroy(ocrb_det *d) {
  if (!d) return OCRB_OK;
  cudaSetDevice(d->ctx->device);
  cudaStreamSynchronize(d->ctx->stream);
  d->stem_w.release(); d->stem_scale.release(); d->stem_shift.release();
  for (auto &kv : d->conv) { kv.second.w32.release(); kv.second.w16.release(); kv.second.scale.release(); kv.second.shift.release(); }
  d->tr1_w32.release(); d->tr1_scale.release(); d->tr1_shift.release(); d->head_w16.release(); d->tr2_w.release();
  for (auto &kv : d->act) kv.second.release();
  d->staged_in.release(); d->staged_out.release(); d->err.release();
  delete d;
  return OCRB_OK;
}

static int det_forward_chunks(ocrb_det *det, const void *images, int dtype, int B, int H, int W, float *prob) {
  ocrb_ctx *ctx = det->ctx;
  const size_t esz = dtype == OCRB_U8 ? 1 : 4;
  const int64_t HW = (int64_t)H * W;
  const int chunk = det->mode == OCRB_MODE_BF16 ? 64 : 16;  // FP32 mode keeps fp32 activations (~0.4 GB per image)
  const bool in_dev = is_device_ptr(images), out_dev = is_device_ptr(prob);
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int bc = B - b0 < chunk ? B - b0 : chunk;
    const uint8_t *src = (const uint8_t *)images + (size_t)b0 * HW * esz;
    float *dst = prob + (size_t)b0 * HW;
    const void *src_dev = src;
    float *dst_dev = dst;
    if (!in_dev) {
      OCRB_TRY(det->staged_in.reserve((size_t)bc * HW * esz));
      OCRB_CUDA(cudaMemcpyAsync(det->staged_in.p, src, (size_t)bc * HW * esz, cudaMemcpyHostToDevice, ctx->stream));
      src_dev = det->staged_in.p;
    }
    if (!out_dev) {
      OCRB_TRY(det->staged_out.reserve((size_t)bc * HW * 4));
      dst_dev = det->staged_out.as<float>();
    }
    OCRB_TRY(det_forward_device(det, src_dev, dtype, bc, H, W, dst_dev, nullptr, 0.6f));
    if (!out_dev) OCRB_CUDA(cudaMemcpyAsync(dst, dst_dev, (size_t)bc * HW * 4, cudaMemcpyDeviceToHost, ctx->stream));
    // host staging buffers are reused by the next chunk
    if (!in_dev || !out_dev) OCRB_TRY(sync(ctx));
  }
  return sync(ctx);
}

int ocrb_det_forward(ocrb_det *det, const void *images, int dtype, int B, int H, int W, float *prob) {
  OCRB_REQUIRE(det && images && prob, "null argument");
  OCRB_REQUIRE(dtype == OCRB_U8 || dtype == OCRB_F32, "unknown dtype %d", dtype);
  OCRB_REQUIRE(B > 0 && H > 0 && W > 0 && H % 32 == 0 && W % 32 == 0, "H and W must be positive multiples of 32 (got %dx%d, B=%d)", H, W, B);
  OCRB_CUDA(cudaSetDevice(det->ctx->device));
  const int rc = det_forward_chunks(det, images, dtype, B, H, W, prob);
  // a bounded pipeline wait that expired traps the kernel: name the wait site instead of the generic CUDA error
  if (det_check_err(det) != OCRB_OK) return OCRB_ERR_INTERNAL;
  return rc;
}

int ocrb_det_tap(ocrb_det *det, const char *name, float *out, int64_t numel) {
  OCRB_REQUIRE(det && name && out, "null argument");
  ocrb_ctx *ctx = det->ctx;
  OCRB_CUDA(cudaSetDevice(ctx->device));
  const int B = det->last_B, H4 = det->last_H / 4, W4 = det->last_W / 4;
  OCRB_REQUIRE(B > 0, "no forward has run yet");
  struct T { const char *tap, *key; int c, div; };
  static const T table[] = {{"stem", "stem", 64, 1}, {"x1", "layer1.1.y", 64, 1}, {"x2", "layer2.1.y", 128, 2},
                            {"x3", "layer3.1.y", 256, 4}, {"x4", "layer4.1.y", 512, 8}, {"fuse", "fuse", 256, 1},
                            {"bin1", "bin1", 64, 1}};
  for (const T &t : table) {
    if (strcmp(t.tap, name) != 0) continue;
    const int h = H4 / t.div, w = W4 / t.div;
    const int64_t n = (int64_t)B * t.c * h * w;
    OCRB_REQUIRE(n == numel, "tap %s has %lld elements, caller passed %lld", name, (long long)n, (long long)numel);
    std::string key = std::string(det->mode == OCRB_MODE_BF16 ? "b." : "f.") + t.key;
    auto it = det->act.find(key);
    OCRB_REQUIRE(it != det->act.end(), "tap %s not available", name);
    DevBuf tmp;
    OCRB_TRY(tmp.reserve((size_t)n * 4));
    if (det->mode == OCRB_MODE_BF16 && det->fpn2_fused && strcmp(name, "fuse") == 0) {
      auto ip = det->act.find("b.cat3");
      OCRB_REQUIRE(ip != det->act.end(), "tap fuse not available");
      fuse_tap_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(it->second.as<__nv_bfloat16>(), ip->second.as<__nv_bfloat16>(), B, h, w, tmp.as<float>());
      OCRB_TRY(check_launch(ctx, "fuse_tap"));
    } else if (det->mode == OCRB_MODE_BF16) {
      bf16_nhwc_to_nchw_f32_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(it->second.as<__nv_bfloat16>(), B, h, w, t.c, t.c, tmp.as<float>());
      OCRB_TRY(check_launch(ctx, "bf16_nhwc_to_nchw_f32"));
    } else {
      OCRB_TRY(launch_nhwc_to_nchw_fp32(ctx, it->second.as<float>(), B, h, w, t.c, t.c, tmp.as<float>()));
    }
    OCRB_CUDA(cudaMemcpyAsync(out, tmp.p, (size_t)n * 4, cudaMemcpyDefault, ctx->stream));
    int rc = sync(ctx);
    tmp.release();
    return rc;
  }
  set_error("unknown tap %s", name);
  return OCRB_ERR_INVALID;
}

}  // extern "C"
