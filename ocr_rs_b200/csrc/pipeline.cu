// run_text_detection's device part for a batch (text_detection/mod.rs:46-67; the batched form
// is get_model_accuracy, :188-204): u8 images -> detector -> binarize -> post-processing,
// plus glyph recognition (char_recognition/mod.rs:39-68) of a caller-provided crop set in the
// same call (the reference has no polygon -> crop glue).
//
// Three streams per context keep the GPU busy across the post-processing's host round trips:
//   copy stream     H2D of image chunk c+1 (double-buffered) while chunk c is in the detector
//   forward stream  detector forward, chunk by chunk (up to 256 images), into a GROUP buffer
//   ctx->stream     post-processing of group g (up to 256 images per launch: the contour /
//                   polygon kernels are latency-bound, so they are amortised over more images)
//                   while the forward of group g+1 runs on the forward stream
#include "common.cuh"

namespace ocrb {

int det_forward_device(ocrb_det *, const void *, int, int, int, int, float *, uint8_t *, float);
ocrb_ctx *det_ctx(ocrb_det *);
int det_mode(ocrb_det *);
int det_check_err(ocrb_det *);
int rec_forward_device(ocrb_rec *, const void *, int, int, float *, int32_t *, double *);
int postproc_device(ocrb_ctx *, const float *, const uint8_t *, const double *, int, int, int, const ocrb_postproc_params &, ocrb_polygons *);
void polygons_append(ocrb_polygons *, const ocrb_polygons *);
ocrb_polygons *polygons_new();
int launch_binarize(ocrb_ctx *, const float *, int64_t, float, uint8_t *);

struct PipelineWorkspace {
  DevBuf images[2], prob[2], bitmap[2], adjust, glyphs, argmax;
  cudaStream_t fwd = nullptr, copy = nullptr;
  cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr}, fwd_done[2] = {nullptr, nullptr}, pp_ready = nullptr;
  bool ready = false;
};
// the workspace belongs to the ctx (one ctx per host thread: no state is shared between contexts)
static int get_ws(ocrb_ctx *ctx, PipelineWorkspace **out) {
  if (!ctx->pipe) ctx->pipe = new PipelineWorkspace();
  PipelineWorkspace *w = ctx->pipe;
  if (!w->ready) {
    OCRB_CUDA(cudaStreamCreateWithFlags(&w->fwd, cudaStreamNonBlocking));
    OCRB_CUDA(cudaStreamCreateWithFlags(&w->copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      OCRB_CUDA(cudaEventCreateWithFlags(&w->copied[i], cudaEventDisableTiming));
      OCRB_CUDA(cudaEventCreateWithFlags(&w->consumed[i], cudaEventDisableTiming));
      OCRB_CUDA(cudaEventCreateWithFlags(&w->fwd_done[i], cudaEventDisableTiming));
    }
    OCRB_CUDA(cudaEventCreateWithFlags(&w->pp_ready, cudaEventDisableTiming));
    w->ready = true;
  }
  *out = w;
  return OCRB_OK;
}

void free_pipe(ocrb_ctx *ctx) {
  PipelineWorkspace *w = ctx->pipe;
  if (!w) return;
  DevBuf *bufs[] = {&w->images[0], &w->images[1], &w->prob[0], &w->prob[1], &w->bitmap[0], &w->bitmap[1], &w->adjust, &w->glyphs, &w->argmax};
  for (DevBuf *b : bufs) b->release();
  for (int i = 0; i < 2; ++i) {
    if (w->copied[i]) cudaEventDestroy(w->copied[i]);
    if (w->consumed[i]) cudaEventDestroy(w->consumed[i]);
    if (w->fwd_done[i]) cudaEventDestroy(w->fwd_done[i]);
  }
  if (w->pp_ready) cudaEventDestroy(w->pp_ready);
  if (w->fwd) cudaStreamDestroy(w->fwd);
  if (w->copy) cudaStreamDestroy(w->copy);
  delete w;
  ctx->pipe = nullptr;
}

constexpr int PIPE_CHUNK_BF16 = 256, PIPE_CHUNK_FP32 = 4, PIPE_GROUP = 256;  // 256 / 256 measured +3 % over 128 / 128 (fewer, longer launches)

}  // namespace ocrb

using namespace ocrb;

extern "C" int ocrb_detect_and_recognize(ocrb_det *det, ocrb_rec *rec, const uint8_t *images, const double *adjust, int B, int H,
                                         int W, const ocrb_postproc_params *params, const uint8_t *glyphs, int n_glyphs,
                                         int32_t *glyph_argmax, ocrb_polygons **out) {
  OCRB_REQUIRE(det && images && adjust && out, "null argument");
  OCRB_REQUIRE(B > 0 && H > 0 && W > 0 && H % 32 == 0 && W % 32 == 0, "H and W must be positive multiples of 32 (got %dx%d, B=%d)", H, W, B);
  OCRB_REQUIRE(n_glyphs == 0 || (rec && glyphs), "glyphs given without a recognition net");
  ocrb_ctx *ctx = det_ctx(det);
  OCRB_CUDA(cudaSetDevice(ctx->device));
  ocrb_postproc_params prm;
  ocrb_postproc_default_params(&prm);
  if (params) prm = *params;
  PipelineWorkspace *ws = nullptr;
  OCRB_TRY(get_ws(ctx, &ws));
  const int64_t HW = (int64_t)H * W;
  const bool bf16 = det_mode(det) == OCRB_MODE_BF16;
  // the per-launch event timeline (ocrb_ctx_profile_begin) needs one stream: serialise then
  const bool serial = ctx->prof.on;
  cudaStream_t s_pp = ctx->stream, s_fwd = serial ? ctx->stream : ws->fwd, s_copy = serial ? ctx->stream : ws->copy;
  // group: <= 256 images and < 2^31 pixels (post-processing index arithmetic); chunk: <= 256 images
  static const int group_env = getenv("OCRB_GROUP") ? atoi(getenv("OCRB_GROUP")) : 0;  // tuning knob
  int group = group_env > 0 ? group_env : PIPE_GROUP;
  // a small batch is still cut into two groups so that post-processing overlaps a forward
  if (B < 2 * group && B >= 64) group = ((B + 1) / 2 + 31) / 32 * 32;
  while ((int64_t)group * HW >= ((int64_t)1 << 31) && group > 1) group /= 2;
  if (group > B) group = B;
  static const int chunk_env = getenv("OCRB_CHUNK") ? atoi(getenv("OCRB_CHUNK")) : 0;  // tuning knob
  int chunk = bf16 ? (chunk_env > 0 ? chunk_env : PIPE_CHUNK_BF16) : PIPE_CHUNK_FP32;
  if (chunk > group) chunk = group;
  const int n_groups = (B + group - 1) / group;
  const bool img_dev = is_device_ptr(images);
  for (int i = 0; i < (n_groups > 1 ? 2 : 1); ++i) {
    OCRB_TRY(ws->prob[i].reserve((size_t)group * HW * 4));
    OCRB_TRY(ws->bitmap[i].reserve((size_t)group * HW));
  }
  if (!img_dev)
    for (int i = 0; i < 2; ++i) OCRB_TRY(ws->images[i].reserve((size_t)chunk * HW));
  OCRB_TRY(ws->adjust.reserve((size_t)B * 16));
  OCRB_CUDA(cudaMemcpyAsync(ws->adjust.p, adjust, (size_t)B * 16, cudaMemcpyDefault, s_pp));
  // everything queued so far on ctx->stream (earlier calls) precedes this call's forward work
  OCRB_CUDA(cudaEventRecord(ws->pp_ready, s_pp));
  OCRB_CUDA(cudaStreamWaitEvent(s_fwd, ws->pp_ready, 0));
  OCRB_CUDA(cudaStreamWaitEvent(s_copy, ws->pp_ready, 0));

  // glyph recognition: independent of the detector, queued first on the post-processing stream
  if (n_glyphs > 0) {
    const void *g = glyphs;
    if (!is_device_ptr(glyphs)) {
      OCRB_TRY(ws->glyphs.reserve((size_t)n_glyphs * 784));
      OCRB_CUDA(cudaMemcpyAsync(ws->glyphs.p, glyphs, (size_t)n_glyphs * 784, cudaMemcpyHostToDevice, s_pp));
      g = ws->glyphs.p;
    }
    int32_t *am = glyph_argmax;
    if (glyph_argmax && !is_device_ptr(glyph_argmax)) {
      OCRB_TRY(ws->argmax.reserve((size_t)n_glyphs * 4));
      am = ws->argmax.as<int32_t>();
    }
    OCRB_TRY(rec_forward_device(rec, g, 1, n_glyphs, nullptr, am, nullptr));
    if (glyph_argmax && am != glyph_argmax)
      OCRB_CUDA(cudaMemcpyAsync(glyph_argmax, am, (size_t)n_glyphs * 4, cudaMemcpyDeviceToHost, s_pp));
  }

  int64_t chunk_no = 0;  // global chunk counter (selects the image staging buffer)
  // queues copies + forwards of group g; the detector launches on ctx->stream, so it is
  // pointed at the forward stream for the duration
  auto enqueue_forward = [&](int g) -> int {
    const int g0 = g * group, gn = B - g0 < group ? B - g0 : group;
    float *prob = ws->prob[g & 1].as<float>();
    uint8_t *bitmap = ws->bitmap[g & 1].as<uint8_t>();
    int rc = OCRB_OK;
    for (int c0 = 0, bc = 0; c0 < gn && rc == OCRB_OK; c0 += bc, ++chunk_no) {
      bc = gn - c0 < chunk ? gn - c0 : chunk;
      // host images: the very first copy of a call is exposed (nothing to overlap it with), so
      // ramp up — 16 images first, the rest of the chunk while those are in the detector
      if (!img_dev && chunk_no == 0 && bc > 32) bc = 16;
      else if (!img_dev && chunk_no == 1 && c0 == 16 && chunk > 16 && gn - c0 > chunk - 16) bc = chunk - 16;
      const uint8_t *src = images + (size_t)(g0 + c0) * HW;
      const int slot = (int)(chunk_no & 1);
      if (!img_dev) {
        if (chunk_no >= 2) OCRB_CUDA(cudaStreamWaitEvent(s_copy, ws->consumed[slot], 0));
        OCRB_CUDA(cudaMemcpyAsync(ws->images[slot].p, src, (size_t)bc * HW, cudaMemcpyHostToDevice, s_copy));
        OCRB_CUDA(cudaEventRecord(ws->copied[slot], s_copy));
        OCRB_CUDA(cudaStreamWaitEvent(s_fwd, ws->copied[slot], 0));
        src = ws->images[slot].as<uint8_t>();
      }
      cudaStream_t saved = ctx->stream;
      ctx->stream = s_fwd;
      rc = det_forward_device(det, src, OCRB_U8, bc, H, W, prob + (size_t)c0 * HW, bf16 ? bitmap + (size_t)c0 * HW : nullptr, (float)prm.thresh);
      if (rc == OCRB_OK && !bf16) rc = launch_binarize(ctx, prob + (size_t)c0 * HW, (int64_t)bc * HW, (float)prm.thresh, bitmap + (size_t)c0 * HW);
      ctx->stream = saved;
      if (rc == OCRB_OK && !img_dev) OCRB_CUDA(cudaEventRecord(ws->consumed[slot], s_fwd));
    }
    if (rc == OCRB_OK) OCRB_CUDA(cudaEventRecord(ws->fwd_done[g & 1], s_fwd));
    return rc;
  };

  ocrb_polygons *res = polygons_new();
  int rc = enqueue_forward(0);
  for (int g = 0; g < n_groups && rc == OCRB_OK; ++g) {
    if (g + 1 < n_groups) rc = enqueue_forward(g + 1);
    if (rc != OCRB_OK) break;
    const int g0 = g * group, gn = B - g0 < group ? B - g0 : group;
    cudaError_t e = cudaStreamWaitEvent(s_pp, ws->fwd_done[g & 1], 0);
    if (e != cudaSuccess) { set_error("cudaStreamWaitEvent -> %s", cudaGetErrorString(e)); rc = OCRB_ERR_CUDA; break; }
    ocrb_polygons *part = polygons_new();
    rc = postproc_device(ctx, ws->prob[g & 1].as<float>(), ws->bitmap[g & 1].as<uint8_t>(), ws->adjust.as<double>() + (size_t)g0 * 2, gn, H, W, prm, part);
    if (rc == OCRB_OK) polygons_append(res, part);
    ocrb_polygons_free(part);
  }
  if (rc == OCRB_OK) rc = sync(ctx);
  cudaStreamSynchronize(ws->fwd);
  cudaStreamSynchronize(ws->copy);
  if (rc == OCRB_OK && bf16) rc = det_check_err(det);
  if (rc != OCRB_OK) {
    cudaStreamSynchronize(ctx->stream);
    ocrb_polygons_free(res);
    return rc;
  }
  *out = res;
  return OCRB_OK;
}
