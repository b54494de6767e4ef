// C ABI: context, error reporting and the image_ops / binarize / CCL entry points.
// (detector: detector.cu, recognition: rec.cu, post-processing: postproc.cu)
#include <algorithm>

#include "common.cuh"

namespace ocrb {
int debug_conv_geometry(int Ho, int Wo, int mode, int *out);
int debug_pipeline_plan(int B, int H, int W, int bf16, int host_images, int *group_out, int *chunks, int cap, int *n_chunks);

static thread_local std::string g_last_error;

void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

int launch_binarize(ocrb_ctx *, const float *, int64_t, float, uint8_t *);
int launch_u8_to_f32(ocrb_ctx *, const uint8_t *, int64_t, float, float *);
int launch_f32_to_u8(ocrb_ctx *, const float *, int64_t, float, uint8_t *);
int launch_preprocess(ocrb_ctx *, const uint8_t *, int, int, int, int, int, int, uint8_t *, uint8_t *);
int launch_preprocess_batch(ocrb_ctx *, const uint8_t *, const void *, int, int, int, int, uint8_t *, int *);
int launch_preprocess_batch_identity(ocrb_ctx *, const uint8_t *, const void *, int, int, int, uint8_t *);
int preprocess_batch_span_limit();
int preprocess_batch_tile_width();
int ccl_canonical_labels(ocrb_ctx *, const uint8_t *, int, int, int, int *, int *);
void free_pp(ocrb_ctx *);
void free_pipe(ocrb_ctx *);

}  // namespace ocrb

using namespace ocrb;

extern "C" {

int ocrb_version(void) { return OCRB_VERSION; }
const char *ocrb_last_error(void) { return g_last_error.c_str(); }

int ocrb_device_count(int *count) {
  OCRB_REQUIRE(count, "null argument");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *count = 0;
    set_error("cudaGetDeviceCount -> %s", cudaGetErrorString(e));
    return OCRB_ERR_CUDA;
  }
  *count = n;
  return OCRB_OK;
}

int ocrb_ctx_create(int device, ocrb_ctx **out) {
  OCRB_REQUIRE(out, "null argument");
  int n = 0;
  OCRB_TRY(ocrb_device_count(&n));
  if (n <= 0) {
    set_error("no CUDA device: libocrb has no CPU fallback");
    return OCRB_ERR_CUDA;
  }
  OCRB_REQUIRE(device >= 0 && device < n, "device %d out of range (have %d)", device, n);
  OCRB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  OCRB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    return OCRB_ERR_CUDA;
  }
  ocrb_ctx *ctx = new ocrb_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  // experiment knob OCRB_PP_HIGHPRIO=1: the context's own stream (post-processing in the pipeline) at the highest priority,
  // so that its CTAs are placed before those of the forward stream's next persistent kernel
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  const bool pp_high = getenv("OCRB_PP_HIGHPRIO") && atoi(getenv("OCRB_PP_HIGHPRIO")) != 0;
  cudaError_t e = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, pp_high ? prio_hi : 0);
  if (e != cudaSuccess) {
    delete ctx;
    set_error("cudaStreamCreate -> %s", cudaGetErrorString(e));
    return OCRB_ERR_CUDA;
  }
  *out = ctx;
  return OCRB_OK;
}

int ocrb_ctx_destroy(ocrb_ctx *ctx) {
  if (!ctx) return OCRB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  free_pp(ctx);
  free_pipe(ctx);
  for (auto &b : ctx->stage) b.release();
  ctx->ccl_tile_empty.release();
  ctx->ccl_seam_list.release();
  ctx->decode_rgba.release();
  for (auto &b : ctx->pin) b.release();
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return OCRB_OK;
}

int ocrb_ctx_synchronize(ocrb_ctx *ctx) {
  OCRB_REQUIRE(ctx, "null ctx");
  return sync(ctx);
}
void *ocrb_ctx_stream(ocrb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int ocrb_ctx_wait_stream(ocrb_ctx *ctx, void *producer_stream) {
  OCRB_REQUIRE(ctx, "null ctx");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  cudaEvent_t ev;
  OCRB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  cudaError_t e = cudaEventRecord(ev, (cudaStream_t)producer_stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ev, 0);
  cudaEventDestroy(ev);  // released once the wait has been satisfied
  if (e != cudaSuccess) {
    set_error("ocrb_ctx_wait_stream -> %s", cudaGetErrorString(e));
    return OCRB_ERR_CUDA;
  }
  return OCRB_OK;
}
int ocrb_ctx_device(ocrb_ctx *ctx) { return ctx ? ctx->device : -1; }
int64_t ocrb_ctx_launch_count(ocrb_ctx *ctx) { return ctx ? ctx->launches : 0; }

int ocrb_ctx_profile_begin(ocrb_ctx *ctx) {
  OCRB_REQUIRE(ctx, "null ctx");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  OCRB_TRY(sync(ctx));
  ctx->prof.used = 0;
  ctx->prof.on = true;
  prof_mark(ctx, "begin");
  return OCRB_OK;
}

int ocrb_ctx_profile_end(ocrb_ctx *ctx, char *buf, size_t cap, size_t *needed) {
  OCRB_REQUIRE(ctx && needed, "null argument");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  Profiler &pr = ctx->prof;
  pr.on = false;
  OCRB_TRY(sync(ctx));
  // aggregate by name, keeping first-seen order
  std::vector<std::string> order;
  std::vector<double> total;
  std::vector<long long> count;
  for (size_t i = 1; i < pr.used; ++i) {
    float ms = 0.f;
    OCRB_CUDA(cudaEventElapsedTime(&ms, pr.ev[i - 1], pr.ev[i]));
    size_t k = 0;
    for (; k < order.size(); ++k)
      if (order[k] == pr.names[i]) break;
    if (k == order.size()) { order.push_back(pr.names[i]); total.push_back(0.0); count.push_back(0); }
    total[k] += ms;
    count[k] += 1;
  }
  std::string out;
  char line[256];
  for (size_t k = 0; k < order.size(); ++k) {
    snprintf(line, sizeof(line), "%s %lld %.6f\n", order[k].c_str(), count[k], total[k]);
    out += line;
  }
  *needed = out.size() + 1;
  if (buf && cap >= out.size() + 1) memcpy(buf, out.c_str(), out.size() + 1);
  return OCRB_OK;
}

// ---- image_ops -------------------------------------------------------------------------
int ocrb_resize_dims(int sw, int sh, int W, int H, int *rw, int *rh) {
  OCRB_REQUIRE(rw && rh && sw > 0 && sh > 0 && W > 0 && H > 0, "bad argument");
  // image 0.23.11 resize_dimensions(fill = false): integer arithmetic
  uint64_t ratio = (uint64_t)sw * (uint64_t)H, nratio = (uint64_t)W * (uint64_t)sh;
  bool use_width = nratio <= ratio;
  uint64_t inter = use_width ? (uint64_t)sh * (uint64_t)W / (uint64_t)sw : (uint64_t)sw * (uint64_t)H / (uint64_t)sh;
  if (inter < 1) inter = 1;
  if (use_width) { *rw = W; *rh = (int)inter; } else { *rw = (int)inter; *rh = H; }
  return OCRB_OK;
}

int ocrb_preprocess_rgba(ocrb_ctx *ctx, const uint8_t *rgba, int sw, int sh, int W, int H, uint8_t *out_gray,
                         double *adjust_x, double *adjust_y) {
  OCRB_REQUIRE(ctx && rgba && out_gray && adjust_x && adjust_y, "null argument");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  int rw, rh;
  OCRB_TRY(ocrb_resize_dims(sw, sh, W, H, &rw, &rh));
  *adjust_x = (double)rw / (double)sw;  // image_ops.rs:201-202
  *adjust_y = (double)rh / (double)sh;
  const void *src = nullptr;
  void *dst = nullptr;
  OCRB_TRY(to_device(ctx, 0, rgba, (size_t)sw * sh * 4, &src));
  OCRB_TRY(out_device(ctx, 1, out_gray, (size_t)W * H, &dst));
  OCRB_TRY(ctx->stage[2].reserve((size_t)sw * rh * 4));
  OCRB_TRY(launch_preprocess(ctx, (const uint8_t *)src, sw, sh, rw, rh, W, H, ctx->stage[2].as<uint8_t>(), (uint8_t *)dst));
  OCRB_TRY(finish_output(ctx, out_gray, dst, (size_t)W * H));
  return sync(ctx);
}

int ocrb_preprocess_rgba_batch(ocrb_ctx *ctx, const uint8_t *rgba, const int64_t *src_offsets, const int *src_w, const int *src_h, int n,
                               int W, int H, uint8_t *out_gray, double *adjust) {
  OCRB_REQUIRE(ctx && rgba && src_offsets && src_w && src_h && out_gray && adjust && n > 0 && W > 0 && H > 0, "bad argument");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  struct PreImageH { int64_t src_off; int sw, sh, rw, rh; };
  std::vector<PreImageH> im((size_t)n);
  int64_t total = 0;
  for (int i = 0; i < n; ++i) {
    OCRB_REQUIRE(src_w[i] > 0 && src_h[i] > 0 && src_offsets[i] >= 0 && src_offsets[i] % 4 == 0, "image %d: bad size or offset", i);
    im[i].src_off = src_offsets[i];
    im[i].sw = src_w[i];
    im[i].sh = src_h[i];
    OCRB_TRY(ocrb_resize_dims(src_w[i], src_h[i], W, H, &im[i].rw, &im[i].rh));
    adjust[2 * i] = (double)im[i].rw / (double)src_w[i];  // image_ops.rs:201-202
    adjust[2 * i + 1] = (double)im[i].rh / (double)src_h[i];
    const int64_t end = src_offsets[i] + (int64_t)src_w[i] * src_h[i] * 4;
    total = end > total ? end : total;
  }
  const void *src = nullptr, *desc = nullptr;
  void *dst = nullptr;
  OCRB_TRY(to_device(ctx, 0, rgba, (size_t)total, &src));
  OCRB_TRY(to_device(ctx, 3, im.data(), im.size() * sizeof(PreImageH), &desc));
  OCRB_TRY(out_device(ctx, 1, out_gray, (size_t)n * W * H, &dst));
  OCRB_TRY(ctx->stage[4].reserve(4));
  OCRB_CUDA(cudaMemsetAsync(ctx->stage[4].p, 0, 4, ctx->stream));
  // shared memory of the fused kernel = the source-column span of one output tile at the batch's largest horizontal
  // down-scaling factor (+ the filter support on either side)
  double worst = 1.0;
  for (int i = 0; i < n; ++i) worst = std::max(worst, (double)im[i].sw / (double)im[i].rw);
  const int span_cap = (int)(preprocess_batch_tile_width() * worst + 2.0 * worst + 8.0);
  int overflow = span_cap > preprocess_batch_span_limit();
  bool all_identity = W % 4 == 0;
  for (int i = 0; i < n; ++i) all_identity = all_identity && im[i].rw == im[i].sw && im[i].rh == im[i].sh;
  if (all_identity) {
    OCRB_TRY(launch_preprocess_batch_identity(ctx, (const uint8_t *)src, desc, n, W, H, (uint8_t *)dst));
    overflow = 0;
  } else if (!overflow) {
    OCRB_TRY(launch_preprocess_batch(ctx, (const uint8_t *)src, desc, n, W, H, span_cap, (uint8_t *)dst, ctx->stage[4].as<int>()));
    OCRB_CUDA(cudaMemcpyAsync(&overflow, ctx->stage[4].p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    OCRB_TRY(sync(ctx));
  }
  if (overflow) {
    // some image is scaled down by more than the fused kernel's tile span allows: the two-kernel path per image
    for (int i = 0; i < n; ++i) {
      OCRB_TRY(ctx->stage[2].reserve((size_t)im[i].sw * im[i].rh * 4));
      OCRB_TRY(launch_preprocess(ctx, (const uint8_t *)src + im[i].src_off, im[i].sw, im[i].sh, im[i].rw, im[i].rh, W, H, ctx->stage[2].as<uint8_t>(),
                                 (uint8_t *)dst + (size_t)i * W * H));
    }
  }
  OCRB_TRY(finish_output(ctx, out_gray, dst, (size_t)n * W * H));
  return sync(ctx);
}

int ocrb_convert_image_to_tensor(ocrb_ctx *ctx, const uint8_t *image, int64_t n, float *out) {
  OCRB_REQUIRE(ctx && image && out && n >= 0, "bad argument");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  const void *src = nullptr;
  void *dst = nullptr;
  OCRB_TRY(to_device(ctx, 0, image, (size_t)n, &src));
  OCRB_TRY(out_device(ctx, 1, out, (size_t)n * 4, &dst));
  OCRB_TRY(launch_u8_to_f32(ctx, (const uint8_t *)src, n, 1.0f, (float *)dst));
  OCRB_TRY(finish_output(ctx, out, dst, (size_t)n * 4));
  return sync(ctx);
}

int ocrb_load_image_as_tensor(ocrb_ctx *ctx, const uint8_t *luma, int64_t n, float *out) {
  OCRB_REQUIRE(ctx && luma && out && n >= 0, "bad argument");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  const void *src = nullptr;
  void *dst = nullptr;
  OCRB_TRY(to_device(ctx, 0, luma, (size_t)n, &src));
  OCRB_TRY(out_device(ctx, 1, out, (size_t)n * 4, &dst));
  OCRB_TRY(launch_u8_to_f32(ctx, (const uint8_t *)src, n, 255.0f, (float *)dst));
  OCRB_TRY(finish_output(ctx, out, dst, (size_t)n * 4));
  return sync(ctx);
}

int ocrb_convert_tensor_to_image(ocrb_ctx *ctx, const float *tensor, int64_t n, float scale, uint8_t *out) {
  OCRB_REQUIRE(ctx && tensor && out && n >= 0, "bad argument");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  const void *src = nullptr;
  void *dst = nullptr;
  OCRB_TRY(to_device(ctx, 0, tensor, (size_t)n * 4, &src));
  OCRB_TRY(out_device(ctx, 1, out, (size_t)n, &dst));
  OCRB_TRY(launch_f32_to_u8(ctx, (const float *)src, n, scale, (uint8_t *)dst));
  OCRB_TRY(finish_output(ctx, out, dst, (size_t)n));
  return sync(ctx);
}

// ---- metrics::binarize -----------------------------------------------------------------
int ocrb_binarize(ocrb_ctx *ctx, const float *pred, int64_t n, double thresh, uint8_t *out) {
  OCRB_REQUIRE(ctx && pred && out && n >= 0, "bad argument");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  const void *src = nullptr;
  void *dst = nullptr;
  OCRB_TRY(to_device(ctx, 0, pred, (size_t)n * 4, &src));
  OCRB_TRY(out_device(ctx, 1, out, (size_t)n, &dst));
  OCRB_TRY(launch_binarize(ctx, (const float *)src, n, (float)thresh, (uint8_t *)dst));
  OCRB_TRY(finish_output(ctx, out, dst, (size_t)n));
  return sync(ctx);
}

int ocrb_ccl_labels(ocrb_ctx *ctx, const uint8_t *bitmap, int B, int H, int W, int32_t *labels, int32_t *n_components) {
  OCRB_REQUIRE(ctx && bitmap && labels && B > 0 && H > 0 && W > 0, "bad argument");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  const int64_t n = (int64_t)B * H * W;
  OCRB_REQUIRE(n < (int64_t)1 << 31, "B*H*W must be < 2^31");
  const void *src = nullptr;
  void *dst = nullptr;
  OCRB_TRY(to_device(ctx, 0, bitmap, (size_t)n, &src));
  OCRB_TRY(out_device(ctx, 1, labels, (size_t)n * 4, &dst));
  OCRB_TRY(ccl_canonical_labels(ctx, (const uint8_t *)src, B, H, W, (int *)dst, n_components));
  OCRB_TRY(finish_output(ctx, labels, dst, (size_t)n * 4));
  return sync(ctx);
}

char ocrb_class_to_char(int cls) {
  static const char *VALUES = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789";  // utils.rs:7
  return (cls >= 0 && cls < 62) ? VALUES[cls] : '?';
}

int ocrb_debug_conv_geometry(int Ho, int Wo, int mode, int *out) {
  OCRB_REQUIRE(out && Ho > 0 && Wo > 0 && mode >= 0 && mode <= 2, "bad argument");
  return ocrb::debug_conv_geometry(Ho, Wo, mode, out);
}

int ocrb_debug_pipeline_plan(int B, int H, int W, int bf16, int host_images, int *group, int *chunks, int cap, int *n_chunks) {
  OCRB_REQUIRE(group && chunks && n_chunks && B > 0 && H > 0 && W > 0 && cap > 0, "bad argument");
  return ocrb::debug_pipeline_plan(B, H, W, bf16, host_images, group, chunks, cap, n_chunks);
}

}  // extern "C"
