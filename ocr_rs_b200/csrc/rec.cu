// char_recognition::model::Net (char_recognition/model.rs:12-39) + the softmax(-1, Double) /
// topk(1) of run_prediction (char_recognition/mod.rs:53-56, utils.rs:28-43).
//
//   view(-1,1,28,28) -> conv 5x5 (1->32, +bias) -> max_pool 2 -> conv 5x5 (32->64, +bias)
//   -> max_pool 2 -> view(-1,1024) [NCHW flatten: c*16 + y*4 + x] -> fc 1024->512 (+bias)
//   -> ReLU -> (dropout off) -> fc 512->62 (+bias).  NB: no ReLU after the convolutions.
//
// FP32-equivalent arithmetic throughout (the north_star asks for a bit-exact class argmax): conv1+pool1 is one
// fused CUDA-core kernel per glyph; conv2 (+pool2+flatten) and fc1 — 88 % of the FLOPs — run on the tensor cores
// with every operand split into two fp16 numbers (rec_tc.cu: 22 significant bits, fp32 accumulation); fc2 runs on
// the shared fp32 implicit-GEMM kernel (conv_fp32.cu); the f64 softmax/argmax is a small kernel.  OCRB_REC=fp32
// selects the all-CUDA-core fp32 path (conv2 / fc1 through conv_fp32.cu too), kept as a bisecting knob.
#include <map>
#include <string>

#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ocrb {

void rec_tc_pack_conv2(const float *, std::vector<uint16_t> &);
void rec_tc_pack_fc(const float *, int, int, std::vector<uint16_t> &);
int launch_rec_conv2_tc(ocrb_ctx *, const __half *, const uint16_t *, const float *, int, __half *, int *);
int launch_rec_conv1_tc(ocrb_ctx *, const uint8_t *, const float *, const float *, int, __half *, int *);
int launch_rec_fc_tc(ocrb_ctx *, const __half *, const uint16_t *, const float *, int, int, int, int, float *, int *);

int launch_conv_fp32(ocrb_ctx *, const float *, int, int, int, int, const float *, int, int, int, int, const float *,
                     const float *, const float *, int, float *);

// ---------------------------------------------------------------------------------------
// conv1 5x5 (1 -> 32) + bias + max_pool 2x2: one CTA per glyph.
// in: [B][784] u8 (x/255 applied here, image_ops.rs:80-83) or f32; out: [B][12][12][32] f32.
// thread = (pooled pixel, 8-channel group): 144 x 4 = 576 items over 192 threads.
// ---------------------------------------------------------------------------------------
constexpr int RC1_THREADS = 192;

// SPLIT: out is [B][144][64] half, channels 0-31 = hi, 32-63 = lo' of the fp16 split (rec_tc.cu) instead of [B][144][32] float
template <class TIn, bool SPLIT>
__global__ void __launch_bounds__(RC1_THREADS) rec_conv1_pool_kernel(const TIn *__restrict__ in, const float *__restrict__ w /*[25][32]*/,
                                                                      const float *__restrict__ bias, int B, void *__restrict__ out_v) {
  __shared__ float s_img[28 * 28];
  __shared__ __align__(16) float s_w[25 * 32];
  __shared__ float s_b[32];
  for (int i = threadIdx.x; i < 800; i += RC1_THREADS) s_w[i] = w[i];
  if (threadIdx.x < 32) s_b[threadIdx.x] = bias[threadIdx.x];
  // persistent CTAs: the weights are staged once, glyphs are taken round robin
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
  __syncthreads();  // the previous glyph's image has been consumed (and, first time, the weights are visible)
  const TIn *img = in + (int64_t)b * 784;
  for (int i = threadIdx.x; i < 784; i += RC1_THREADS) {
    if (sizeof(TIn) == 1) s_img[i] = (float)img[i] / 255.0f;
    else s_img[i] = (float)img[i];
  }
  __syncthreads();
  for (int item = threadIdx.x; item < 576; item += RC1_THREADS) {
    const int cg = item & 3, pp = item >> 2;
    const int py = pp / 12, px = pp - py * 12;
    float patch[36];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int s = 0; s < 6; ++s) patch[r * 6 + s] = s_img[(2 * py + r) * 28 + 2 * px + s];
    // the four conv positions of the pool window share every weight load; channel pairs go through packed FFMA2
    float2 acc[4][4];
#pragma unroll
    for (int pos = 0; pos < 4; ++pos)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[pos][j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
      for (int s = 0; s < 5; ++s) {
        const float4 w0 = *reinterpret_cast<const float4 *>(&s_w[(r * 5 + s) * 32 + cg * 8]);
        const float4 w1 = *reinterpret_cast<const float4 *>(&s_w[(r * 5 + s) * 32 + cg * 8 + 4]);
#pragma unroll
        for (int pos = 0; pos < 4; ++pos) {
          const float v = patch[((pos >> 1) + r) * 6 + (pos & 1) + s];
          const float2 vv = make_float2(v, v);
          acc[pos][0] = ffma2(vv, make_float2(w0.x, w0.y), acc[pos][0]);
          acc[pos][1] = ffma2(vv, make_float2(w0.z, w0.w), acc[pos][1]);
          acc[pos][2] = ffma2(vv, make_float2(w1.x, w1.y), acc[pos][2]);
          acc[pos][3] = ffma2(vv, make_float2(w1.z, w1.w), acc[pos][3]);
        }
      }
    float best[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      best[2 * j] = fmaxf(fmaxf(acc[0][j].x, acc[1][j].x), fmaxf(acc[2][j].x, acc[3][j].x));
      best[2 * j + 1] = fmaxf(fmaxf(acc[0][j].y, acc[1][j].y), fmaxf(acc[2][j].y, acc[3][j].y));
    }
    // max(conv) + bias == max(conv + bias): rounding is monotone
    if (!SPLIT) {
      float *op = reinterpret_cast<float *>(out_v) + ((int64_t)b * 144 + pp) * 32 + cg * 8;
      *reinterpret_cast<float4 *>(op) = make_float4(best[0] + s_b[cg * 8 + 0], best[1] + s_b[cg * 8 + 1], best[2] + s_b[cg * 8 + 2], best[3] + s_b[cg * 8 + 3]);
      *reinterpret_cast<float4 *>(op + 4) = make_float4(best[4] + s_b[cg * 8 + 4], best[5] + s_b[cg * 8 + 5], best[6] + s_b[cg * 8 + 6], best[7] + s_b[cg * 8 + 7]);
    } else {
      __half hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = best[j] + s_b[cg * 8 + j];
        hi[j] = __float2half_rn(v);
        lo[j] = __float2half_rn((v - __half2float(hi[j])) * 2048.0f);
      }
      __half *op = reinterpret_cast<__half *>(out_v) + ((int64_t)b * 144 + pp) * 64 + cg * 8;
      *reinterpret_cast<uint4 *>(op) = *reinterpret_cast<uint4 *>(hi);
      *reinterpret_cast<uint4 *>(op + 32) = *reinterpret_cast<uint4 *>(lo);
    }
  }
  }
}

// max_pool 2x2 over [B][8][8][64] NHWC + NCHW flatten -> [B][1024] (index c*16 + y*4 + x)
__global__ void rec_pool2_flatten_kernel(const float *__restrict__ in, int64_t B, float *__restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * 1024) return;
  // threads of a warp walk channels (contiguous NHWC reads); the write is strided by 16 floats
  const int c = (int)(idx & 63), pp = (int)((idx >> 6) & 15);
  const int64_t b = idx >> 10;
  const int y = pp >> 2, x = pp & 3;
  const float *ip = in + ((b * 8 + 2 * y) * 8 + 2 * x) * 64 + c;
  const float m = fmaxf(fmaxf(ip[0], ip[64]), fmaxf(ip[8 * 64], ip[8 * 64 + 64]));
  out[b * 1024 + c * 16 + pp] = m;
}

// logits [B][ld] (first 62 valid) -> logits_out [B][62], argmax [B], softmax(-1, Double) top-1 prob [B]
__global__ void rec_top1_kernel(const float *__restrict__ logits, int ld, int64_t B, float *__restrict__ logits_out,
                                int32_t *__restrict__ argmax, double *__restrict__ prob) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float *lp = logits + b * ld;
  float best = lp[0];
  int bi = 0;
  for (int i = 1; i < 62; ++i) {
    const float v = lp[i];
    if (v > best) { best = v; bi = i; }  // first maximum wins
  }
  if (logits_out)
    for (int i = 0; i < 62; ++i) logits_out[b * 62 + i] = lp[i];
  if (argmax) argmax[b] = bi;
  if (prob) {
    double s = 0.0;
    for (int i = 0; i < 62; ++i) s += exp((double)lp[i] - (double)best);
    prob[b] = 1.0 / s;
  }
}

}  // namespace ocrb

using namespace ocrb;

struct ocrb_rec {
  ocrb_ctx *ctx = nullptr;
  DevBuf w1, b1;          // [25][32], [32]
  DevBuf w2, b2, one64;   // [25][32][64], [64]
  DevBuf w3, b3, one512;  // [1024][512], [512]
  DevBuf w4, b4;          // [512][64] (62 padded), [64]
  DevBuf w2s, w3s;        // fp16-split packings for the tensor-core kernels (rec_tc.cu), fc1 bias = b3
  DevBuf a1, a2, a3, a4, a5, in_stage, out_logits, out_arg, out_prob, err;
};

namespace ocrb {

template <class T>
static int upload_vec(DevBuf &buf, const std::vector<T> &v) {
  OCRB_TRY(buf.reserve(v.size() * sizeof(T)));
  OCRB_CUDA(cudaMemcpy(buf.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return OCRB_OK;
}

// device-side forward on already-resident glyphs; outputs may be null
int rec_forward_device(ocrb_rec *r, const void *glyphs_dev, int is_u8, int B, float *logits_dev, int32_t *argmax_dev, double *prob_dev) {
  ocrb_ctx *ctx = r->ctx;
  static const bool fp32_path = getenv("OCRB_REC") && strcmp(getenv("OCRB_REC"), "fp32") == 0;
  const int rc1_grid = B < 8 * ctx->sm_count ? B : 8 * ctx->sm_count;
  OCRB_TRY(r->a4.reserve((size_t)B * 512 * 4));
  OCRB_TRY(r->a5.reserve((size_t)B * 64 * 4));
  if (!fp32_path) {
    // tensor-core path: conv1 (CUDA cores) -> split halves -> conv2 + pool2 + flatten -> fc1 + ReLU (tcgen05, fp16 split)
    OCRB_TRY(r->a1.reserve((size_t)B * 144 * 64 * 2));
    OCRB_TRY(r->a3.reserve((size_t)B * 2048 * 2));
    OCRB_TRY(r->err.reserve(4));
    // conv1: u8 glyphs (16-byte aligned) on the tensor cores (rec_tc.cu); f32 glyphs / OCRB_REC_CONV1=cuda on CUDA cores
    static const bool conv1_cuda = getenv("OCRB_REC_CONV1") && strcmp(getenv("OCRB_REC_CONV1"), "cuda") == 0;
    if (is_u8 && !conv1_cuda && (reinterpret_cast<uintptr_t>(glyphs_dev) & 15) == 0) {
      OCRB_TRY(launch_rec_conv1_tc(ctx, (const uint8_t *)glyphs_dev, r->w1.as<float>(), r->b1.as<float>(), B, r->a1.as<__half>(), r->err.as<int>()));
    } else {
      if (is_u8)
        rec_conv1_pool_kernel<uint8_t, true><<<rc1_grid, RC1_THREADS, 0, ctx->stream>>>((const uint8_t *)glyphs_dev, r->w1.as<float>(), r->b1.as<float>(), B, r->a1.p);
      else
        rec_conv1_pool_kernel<float, true><<<rc1_grid, RC1_THREADS, 0, ctx->stream>>>((const float *)glyphs_dev, r->w1.as<float>(), r->b1.as<float>(), B, r->a1.p);
      OCRB_TRY(check_launch(ctx, "rec_conv1_pool"));
    }
    OCRB_TRY(launch_rec_conv2_tc(ctx, r->a1.as<__half>(), r->w2s.as<uint16_t>(), r->b2.as<float>(), B, r->a3.as<__half>(), r->err.as<int>()));
    OCRB_TRY(launch_rec_fc_tc(ctx, r->a3.as<__half>(), r->w3s.as<uint16_t>(), r->b3.as<float>(), B, 1024, 512, 1, r->a4.as<float>(), r->err.as<int>()));
  } else {
  OCRB_TRY(r->a1.reserve((size_t)B * 144 * 32 * 4));
  OCRB_TRY(r->a2.reserve((size_t)B * 64 * 64 * 4));
  OCRB_TRY(r->a3.reserve((size_t)B * 1024 * 4));
  if (is_u8)
    rec_conv1_pool_kernel<uint8_t, false><<<rc1_grid, RC1_THREADS, 0, ctx->stream>>>((const uint8_t *)glyphs_dev, r->w1.as<float>(), r->b1.as<float>(), B, r->a1.p);
  else
    rec_conv1_pool_kernel<float, false><<<rc1_grid, RC1_THREADS, 0, ctx->stream>>>((const float *)glyphs_dev, r->w1.as<float>(), r->b1.as<float>(), B, r->a1.p);
  OCRB_TRY(check_launch(ctx, "rec_conv1_pool"));
  // conv2 5x5 32 -> 64 on [B][12][12][32] -> [B][8][8][64]
  OCRB_TRY(launch_conv_fp32(ctx, r->a1.as<float>(), B, 12, 12, 32, r->w2.as<float>(), 64, 5, 1, 0, r->one64.as<float>(), r->b2.as<float>(), nullptr, 0, r->a2.as<float>()));
  rec_pool2_flatten_kernel<<<(unsigned)cdiv((int64_t)B * 1024, 256), 256, 0, ctx->stream>>>(r->a2.as<float>(), B, r->a3.as<float>());
  OCRB_TRY(check_launch(ctx, "rec_pool2_flatten"));
  OCRB_TRY(launch_conv_fp32(ctx, r->a3.as<float>(), B, 1, 1, 1024, r->w3.as<float>(), 512, 1, 1, 0, r->one512.as<float>(), r->b3.as<float>(), nullptr, 1, r->a4.as<float>()));
  }
  // fc2 as a 1x1 "convolution" over B pixels
  OCRB_TRY(launch_conv_fp32(ctx, r->a4.as<float>(), B, 1, 1, 512, r->w4.as<float>(), 64, 1, 1, 0, r->one64.as<float>(), r->b4.as<float>(), nullptr, 0, r->a5.as<float>()));
  rec_top1_kernel<<<(unsigned)cdiv(B, 128), 128, 0, ctx->stream>>>(r->a5.as<float>(), 64, B, logits_dev, argmax_dev, prob_dev);
  return check_launch(ctx, "rec_top1");
}

static int rec_forward_any(ocrb_rec *rec, const void *glyphs, int is_u8, int B, float *logits, int32_t *argmax, double *prob) {
  OCRB_REQUIRE(rec && glyphs && B > 0, "bad argument");
  ocrb_ctx *ctx = rec->ctx;
  OCRB_CUDA(cudaSetDevice(ctx->device));
  const size_t in_bytes = (size_t)B * 784 * (is_u8 ? 1 : 4);
  const void *src = glyphs;
  if (!is_device_ptr(glyphs)) {
    OCRB_TRY(rec->in_stage.reserve(in_bytes));
    OCRB_CUDA(cudaMemcpyAsync(rec->in_stage.p, glyphs, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    src = rec->in_stage.p;
  }
  float *lg = logits;
  int32_t *am = argmax;
  double *pr = prob;
  if (logits && !is_device_ptr(logits)) { OCRB_TRY(rec->out_logits.reserve((size_t)B * 62 * 4)); lg = rec->out_logits.as<float>(); }
  if (argmax && !is_device_ptr(argmax)) { OCRB_TRY(rec->out_arg.reserve((size_t)B * 4)); am = rec->out_arg.as<int32_t>(); }
  if (prob && !is_device_ptr(prob)) { OCRB_TRY(rec->out_prob.reserve((size_t)B * 8)); pr = rec->out_prob.as<double>(); }
  OCRB_TRY(rec_forward_device(rec, src, is_u8, B, lg, am, pr));
  if (logits && lg != logits) OCRB_CUDA(cudaMemcpyAsync(logits, lg, (size_t)B * 62 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (argmax && am != argmax) OCRB_CUDA(cudaMemcpyAsync(argmax, am, (size_t)B * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (prob && pr != prob) OCRB_CUDA(cudaMemcpyAsync(prob, pr, (size_t)B * 8, cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}

}  // namespace ocrb

extern "C" {

int ocrb_rec_create(ocrb_ctx *ctx, int n, const char *const *names, const float *const *data, const int64_t *numel, ocrb_rec **out) {
  OCRB_REQUIRE(ctx && names && data && numel && out && n > 0, "bad argument");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  // Canonical names ("conv1.weight" ...), or the VarStore names of the reference's model file.  There all four
  // layers are created on ONE path (char_recognition/model.rs:14-17), so the names collide and tch de-duplicates
  // them with a "__<n>" suffix whose numbering depends on the creation order inside nn::conv / nn::linear
  // (bias-first for conv; linear differs between tch versions).  The suffix is therefore NOT
  // interpreted: a tensor called weight* / bias* is placed by its element count, and the eight counts are distinct.
  static const char *canon[8] = {"conv1.bias", "conv1.weight", "conv2.bias", "conv2.weight", "fc1.bias", "fc1.weight", "fc2.bias", "fc2.weight"};
  static const int64_t sizes[8] = {32, 32 * 25, 64, 64 * 32 * 25, 512, 512 * 1024, 62, 62 * 512};
  const float *t[8] = {nullptr};
  for (int i = 0; i < n; ++i) {
    OCRB_REQUIRE(names[i] && data[i], "bad tensor %d", i);
    int slot = -1;
    for (int k = 0; k < 8; ++k)
      if (strcmp(names[i], canon[k]) == 0) slot = k;
    if (slot < 0) {
      std::string base(names[i]);
      const size_t us = base.find("__");
      if (us != std::string::npos) base.resize(us);
      const size_t dot = base.rfind('.');
      if (dot != std::string::npos) base = base.substr(dot + 1);
      const int kind = base == "bias" ? 0 : (base == "weight" ? 1 : -1);
      if (kind < 0) continue;  // not a tensor of this net
      for (int k = kind; k < 8; k += 2)
        if (numel[i] == sizes[k]) slot = k;
      OCRB_REQUIRE(slot >= 0, "tensor %s has %lld elements: no %s of the glyph net has that size", names[i], (long long)numel[i], base.c_str());
    }
    OCRB_REQUIRE(numel[i] == sizes[slot], "tensor %s has %lld elements, expected %lld", names[i], (long long)numel[i], (long long)sizes[slot]);
    OCRB_REQUIRE(!t[slot], "tensor %s: %s given twice", names[i], canon[slot]);
    t[slot] = data[i];
  }
  for (int k = 0; k < 8; ++k) OCRB_REQUIRE(t[k], "missing tensor %s", canon[k]);
  ocrb_rec *r = new ocrb_rec();
  r->ctx = ctx;
  auto fail = [&](int rc) { ocrb_rec_destroy(r); return rc; };
  int rc;
  {  // conv1 [32][1][5][5] -> [25][32]
    std::vector<float> w(800), b(t[0], t[0] + 32);
    for (int co = 0; co < 32; ++co)
      for (int tp = 0; tp < 25; ++tp) w[tp * 32 + co] = t[1][co * 25 + tp];
    if ((rc = upload_vec(r->w1, w)) || (rc = upload_vec(r->b1, b))) return fail(rc);
  }
  {  // conv2 [64][32][5][5] -> [25][32][64]
    std::vector<float> w((size_t)25 * 32 * 64), b(t[2], t[2] + 64), one(64, 1.0f);
    for (int co = 0; co < 64; ++co)
      for (int ci = 0; ci < 32; ++ci)
        for (int tp = 0; tp < 25; ++tp) w[((size_t)tp * 32 + ci) * 64 + co] = t[3][((size_t)co * 32 + ci) * 25 + tp];
    if ((rc = upload_vec(r->w2, w)) || (rc = upload_vec(r->b2, b)) || (rc = upload_vec(r->one64, one))) return fail(rc);
    std::vector<uint16_t> ws;
    rec_tc_pack_conv2(t[3], ws);
    if ((rc = upload_vec(r->w2s, ws))) return fail(rc);
  }
  {  // fc1 [512][1024] -> [1024][512]
    std::vector<float> w((size_t)1024 * 512), b(t[4], t[4] + 512), one(512, 1.0f);
    for (int o = 0; o < 512; ++o)
      for (int i = 0; i < 1024; ++i) w[(size_t)i * 512 + o] = t[5][(size_t)o * 1024 + i];
    if ((rc = upload_vec(r->w3, w)) || (rc = upload_vec(r->b3, b)) || (rc = upload_vec(r->one512, one))) return fail(rc);
    std::vector<uint16_t> ws;
    rec_tc_pack_fc(t[5], 512, 1024, ws);
    if ((rc = upload_vec(r->w3s, ws))) return fail(rc);
  }
  {  // fc2 [62][512] -> [512][64] zero padded
    std::vector<float> w((size_t)512 * 64, 0.0f), b(64, 0.0f);
    for (int o = 0; o < 62; ++o) {
      b[o] = t[6][o];
      for (int i = 0; i < 512; ++i) w[(size_t)i * 64 + o] = t[7][(size_t)o * 512 + i];
    }
    if ((rc = upload_vec(r->w4, w)) || (rc = upload_vec(r->b4, b))) return fail(rc);
  }
  *out = r;
  return OCRB_OK;
}

int ocrb_rec_destroy(ocrb_rec *r) {
  if (!r) return OCRB_OK;
  cudaSetDevice(r->ctx->device);
  cudaStreamSynchronize(r->ctx->stream);
  DevBuf *all[] = {&r->w1, &r->b1, &r->w2, &r->b2, &r->one64, &r->w3, &r->b3, &r->one512, &r->w4, &r->b4, &r->a1, &r->a2,
                   &r->a3, &r->a4, &r->a5, &r->in_stage, &r->out_logits, &r->out_arg, &r->out_prob, &r->w2s, &r->w3s, &r->err};
  for (DevBuf *b : all) b->release();
  delete r;
  return OCRB_OK;
}

int ocrb_rec_forward(ocrb_rec *rec, const float *glyphs, int B, float *logits, int32_t *argmax, double *prob) {
  return rec_forward_any(rec, glyphs, 0, B, logits, argmax, prob);
}
int ocrb_rec_forward_u8(ocrb_rec *rec, const uint8_t *glyphs, int B, float *logits, int32_t *argmax, double *prob) {
  return rec_forward_any(rec, glyphs, 1, B, logits, argmax, prob);
}

}  // extern "C"
