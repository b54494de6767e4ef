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
//
// OCRB_PP_SMS=k (k a multiple of 8): the device's SMs are split into two green contexts (driver resource partition):
// 148 - k SMs for the forward stream — every convolution kernel is persistent with one CTA per SM of ITS partition — and k
// SMs on which the latency-bound post-processing kernels of a group run BESIDE the forward of the next group.  Without the
// partition the two streams never overlap: a persistent convolution kernel fills every SM, so a post-processing kernel
// only gets SMs at a kernel boundary and then holds the next convolution kernel's CTAs off them.  The glyph net of such a
// group runs on the forward partition (tensor-core kernels), the last group's post-processing on the whole device.
#include <cuda.h>

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
void postproc_kept_boxes(ocrb_ctx *, const int2 **, const int **);
int launch_crop_glyphs(ocrb_ctx *, const uint8_t *, int, int, const int2 *, const int *, int, int, uint8_t *);

struct PipelineWorkspace {
  DevBuf images[2], prob[2], bitmap[2], adjust, glyphs, argmax;  // images: one staging buffer per GROUP (kept until its crops are cut)
  DevBuf crops[2], crop_cls[2];                                   // glyph tiles / classes of a group's kept polygons
  std::vector<PinBuf> cls_host;                                   // per group: classes on the host, read after the final sync
  cudaStream_t fwd = nullptr, copy = nullptr;
  cudaEvent_t copied[2] = {nullptr, nullptr}, img_free[2] = {nullptr, nullptr}, fwd_done[2] = {nullptr, nullptr}, pp_ready = nullptr;
  // SM partition (OCRB_PP_SMS): green contexts, the post-processing stream of the small one, SMs of the large one
  CUgreenCtx g_fwd = nullptr, g_pp = nullptr;
  cudaStream_t pp_small = nullptr;
  int fwd_sms = 0;
  cudaEvent_t pp_chain = nullptr, crop_done = nullptr, rec_done[2] = {nullptr, nullptr};
  bool ready = false;
};

// driver entry points of the green-context API, resolved at run time (the library does not link libcuda)
template <class F>
static bool drv_fn(const char *name, F *fn) {
  void *ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !ptr) {
    cudaGetLastError();
    return false;
  }
  *fn = reinterpret_cast<F>(ptr);
  return true;
}

// splits the device into (sm_count - pp_sms) + pp_sms SMs; on any failure the pipeline keeps its plain streams
static bool make_partition(ocrb_ctx *ctx, PipelineWorkspace *w, int pp_sms) {
  CUresult (*getRes)(CUdevice, CUdevResource *, CUdevResourceType) = nullptr;
  CUresult (*split)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int) = nullptr;
  CUresult (*genDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int) = nullptr;
  CUresult (*gcCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
  CUresult (*gcStream)(CUstream *, CUgreenCtx, unsigned int, int) = nullptr;
  if (!drv_fn("cuDeviceGetDevResource", &getRes) || !drv_fn("cuDevSmResourceSplitByCount", &split) || !drv_fn("cuDevResourceGenerateDesc", &genDesc) ||
      !drv_fn("cuGreenCtxCreate", &gcCreate) || !drv_fn("cuGreenCtxStreamCreate", &gcStream))
    return false;
  CUdevResource all, small, rest;
  unsigned int groups = 1;
  if (getRes((CUdevice)ctx->device, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
  if (split(&small, &groups, &all, &rest, 0, (unsigned int)pp_sms) != CUDA_SUCCESS || groups != 1) return false;
  if (small.sm.smCount < 1 || rest.sm.smCount < 64) return false;
  CUdevResourceDesc d_small, d_rest;
  if (genDesc(&d_small, &small, 1) != CUDA_SUCCESS || genDesc(&d_rest, &rest, 1) != CUDA_SUCCESS) return false;
  if (gcCreate(&w->g_pp, d_small, (CUdevice)ctx->device, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
  if (gcCreate(&w->g_fwd, d_rest, (CUdevice)ctx->device, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
  CUstream s_small = nullptr, s_fwd = nullptr;
  if (gcStream(&s_small, w->g_pp, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return false;
  if (gcStream(&s_fwd, w->g_fwd, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return false;
  w->pp_small = s_small;
  w->fwd = s_fwd;
  w->fwd_sms = (int)rest.sm.smCount;
  return true;
}
// the workspace belongs to the ctx (one ctx per host thread: no state is shared between contexts)
static int get_ws(ocrb_ctx *ctx, PipelineWorkspace **out) {
  if (!ctx->pipe) ctx->pipe = new PipelineWorkspace();
  PipelineWorkspace *w = ctx->pipe;
  if (!w->ready) {
    static const int pp_sms = getenv("OCRB_PP_SMS") ? atoi(getenv("OCRB_PP_SMS")) : 0;
    if (pp_sms > 0 && !make_partition(ctx, w, pp_sms)) {
      fprintf(stderr, "libocrb: OCRB_PP_SMS=%d: no green-context partition on this driver / device, plain streams\n", pp_sms);
      w->pp_small = nullptr;
      w->fwd = nullptr;
      w->fwd_sms = 0;
    }
    if (!w->fwd) OCRB_CUDA(cudaStreamCreateWithFlags(&w->fwd, cudaStreamNonBlocking));
    OCRB_CUDA(cudaStreamCreateWithFlags(&w->copy, cudaStreamNonBlocking));
    OCRB_CUDA(cudaEventCreateWithFlags(&w->pp_chain, cudaEventDisableTiming));
    OCRB_CUDA(cudaEventCreateWithFlags(&w->crop_done, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) OCRB_CUDA(cudaEventCreateWithFlags(&w->rec_done[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
      OCRB_CUDA(cudaEventCreateWithFlags(&w->copied[i], cudaEventDisableTiming));
      OCRB_CUDA(cudaEventCreateWithFlags(&w->img_free[i], cudaEventDisableTiming));
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
  DevBuf *bufs[] = {&w->images[0], &w->images[1], &w->prob[0], &w->prob[1], &w->bitmap[0], &w->bitmap[1], &w->adjust, &w->glyphs, &w->argmax,
                    &w->crops[0], &w->crops[1], &w->crop_cls[0], &w->crop_cls[1]};
  for (DevBuf *b : bufs) b->release();
  for (PinBuf &b : w->cls_host) b.release();
  for (int i = 0; i < 2; ++i) {
    if (w->copied[i]) cudaEventDestroy(w->copied[i]);
    if (w->img_free[i]) cudaEventDestroy(w->img_free[i]);
    if (w->fwd_done[i]) cudaEventDestroy(w->fwd_done[i]);
  }
  if (w->pp_ready) cudaEventDestroy(w->pp_ready);
  if (w->pp_chain) cudaEventDestroy(w->pp_chain);
  if (w->crop_done) cudaEventDestroy(w->crop_done);
  for (int i = 0; i < 2; ++i)
    if (w->rec_done[i]) cudaEventDestroy(w->rec_done[i]);
  if (w->pp_small) cudaStreamDestroy(w->pp_small);
  if (w->fwd) cudaStreamDestroy(w->fwd);
  if (w->copy) cudaStreamDestroy(w->copy);
  if (w->g_fwd || w->g_pp) {
    CUresult (*gcDestroy)(CUgreenCtx) = nullptr;
    if (drv_fn("cuGreenCtxDestroy", &gcDestroy)) {
      if (w->g_fwd) gcDestroy(w->g_fwd);
      if (w->g_pp) gcDestroy(w->g_pp);
    }
  }
  delete w;
  ctx->pipe = nullptr;
}

void polygons_set_glyph_classes(ocrb_polygons *, int, const std::vector<PinBuf> &, const std::vector<int64_t> &);

constexpr int PIPE_CHUNK_BF16 = 256, PIPE_CHUNK_FP32 = 16, PIPE_GROUP = 256;  // 256 / 256 measured +3 % over 128 / 128 (fewer, longer launches)

// The batching plan of a call: images per post-processing group and per forward chunk.
// group: <= 256 images and < 2^31 pixels (post-processing index arithmetic); chunk: <= 256 images, never across a group
static void pipeline_plan(int B, int64_t HW, bool bf16, int *group_out, int *chunk_out) {
  static const int group_env = getenv("OCRB_GROUP") ? atoi(getenv("OCRB_GROUP")) : 0;  // tuning knob
  int group = group_env > 0 ? group_env : PIPE_GROUP;
  // OCRB_GROUP_SPLIT=1: a small batch is still cut into two groups (round 1's rule, "so that post-processing overlaps a
  // forward").  Off by default: the streams hide host round trips, not SM time, and one longer forward is more efficient —
  // measured 9.99 against 10.48 ms per 128 images, 19.4 against 20.0 per 256 (the per-rank batches at 8 and 4 GPUs)
  static const bool split_small = getenv("OCRB_GROUP_SPLIT") && atoi(getenv("OCRB_GROUP_SPLIT")) == 1;  // tuning knob
  if (split_small && B < 2 * group && B >= 64) group = ((B + 1) / 2 + 31) / 32 * 32;
  while ((int64_t)group * HW >= ((int64_t)1 << 31) && group > 1) group /= 2;
  if (group > B) group = B;
  static const int chunk_env = getenv("OCRB_CHUNK") ? atoi(getenv("OCRB_CHUNK")) : 0;  // tuning knob
  int chunk = bf16 ? (chunk_env > 0 ? chunk_env : PIPE_CHUNK_BF16) : PIPE_CHUNK_FP32;
  if (chunk > group) chunk = group;
  *group_out = group;
  *chunk_out = chunk;
}

// Images of the next forward chunk when `remaining` images of the group are left.  Host images: the very first copy of a
// call is exposed (nothing to overlap it with), so ramp up — 16 images first, then chunks three times the previous one (a
// chunk's copy takes about a third of its forward when several GPUs share the host's memory and PCIe switches, so every
// copy hides behind the forward before it).  *ramp starts at 16 per call.
static int pipeline_next_chunk(int remaining, int chunk, bool host_images, int *ramp) {
  int bc = remaining < chunk ? remaining : chunk;
  if (host_images && *ramp < chunk) {
    if (bc > *ramp) bc = *ramp;
    *ramp *= 3;
  }
  return bc;
}

// host-only view of the plan for tests (include/ocrb.h ocrb_debug_pipeline_plan)
int debug_pipeline_plan(int B, int H, int W, int bf16, int host_images, int *group_out, int *chunks, int cap, int *n_chunks) {
  int group = 0, chunk = 0;
  pipeline_plan(B, (int64_t)H * W, bf16 != 0, &group, &chunk);
  int n = 0, ramp = 16;
  for (int g0 = 0; g0 < B; g0 += group) {
    const int gn = B - g0 < group ? B - g0 : group;
    for (int c0 = 0, bc = 0; c0 < gn; c0 += bc) {
      bc = pipeline_next_chunk(gn - c0, chunk, host_images != 0, &ramp);
      if (n < cap) chunks[n] = bc;
      ++n;
    }
  }
  *group_out = group;
  *n_chunks = n;
  return n <= cap ? OCRB_OK : OCRB_ERR_CAPACITY;
}

}  // namespace ocrb

using namespace ocrb;

// crop_k > 0: the polygon -> glyph crop glue (crop.cu) feeds the recognition net from the detector's own polygons;
// crop_k == 0: the caller's glyph set is classified (ocrb_detect_and_recognize).
static int run_pipeline(ocrb_det *det, ocrb_rec *rec, const uint8_t *images, const double *adjust, int B, int H, int W,
                        const ocrb_postproc_params *params, const uint8_t *glyphs, int n_glyphs, int32_t *glyph_argmax, int crop_k,
                        ocrb_polygons **out) {
  OCRB_REQUIRE(det && images && adjust && out, "null argument");
  OCRB_REQUIRE(B > 0 && H > 0 && W > 0 && H % 32 == 0 && W % 32 == 0, "H and W must be positive multiples of 32 (got %dx%d, B=%d)", H, W, B);
  OCRB_REQUIRE(n_glyphs == 0 || (rec && glyphs), "glyphs given without a recognition net");
  OCRB_REQUIRE(crop_k == 0 || rec, "glyph crops asked for without a recognition net");
  OCRB_REQUIRE(crop_k >= 0 && crop_k <= 64, "glyphs_per_polygon must be in 0..64 (got %d)", crop_k);
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
  const bool split = ws->pp_small != nullptr && !serial;  // SM partition in use
  int group = 0, chunk = 0;
  pipeline_plan(B, HW, bf16, &group, &chunk);
  const int n_groups = (B + group - 1) / group;
  const bool img_dev = is_device_ptr(images);
  for (int i = 0; i < (n_groups > 1 ? 2 : 1); ++i) {
    OCRB_TRY(ws->prob[i].reserve((size_t)group * HW * 4));
    OCRB_TRY(ws->bitmap[i].reserve((size_t)group * HW));
    if (!img_dev) OCRB_TRY(ws->images[i].reserve((size_t)group * HW));
  }
  if (crop_k > 0 && (int)ws->cls_host.size() < n_groups) ws->cls_host.resize(n_groups);
  OCRB_TRY(ws->adjust.reserve((size_t)B * 16));
  OCRB_CUDA(cudaMemcpyAsync(ws->adjust.p, adjust, (size_t)B * 16, cudaMemcpyDefault, s_pp));
  // everything queued so far on ctx->stream (earlier calls) precedes this call's forward work
  OCRB_CUDA(cudaEventRecord(ws->pp_ready, s_pp));
  OCRB_CUDA(cudaStreamWaitEvent(s_fwd, ws->pp_ready, 0));
  OCRB_CUDA(cudaStreamWaitEvent(s_copy, ws->pp_ready, 0));

  // caller-provided glyphs: independent of the detector, queued first on the post-processing stream
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

  int ramp = 16;  // images of the next ramp-up chunk (host images only)
  // queues copies + forwards of group g; the detector launches on ctx->stream, so it is
  // pointed at the forward stream for the duration
  auto enqueue_forward = [&](int g) -> int {
    const int g0 = g * group, gn = B - g0 < group ? B - g0 : group;
    float *prob = ws->prob[g & 1].as<float>();
    uint8_t *bitmap = ws->bitmap[g & 1].as<uint8_t>();
    int rc = OCRB_OK;
    // the group's staging buffer is free once group g - 2 has been forwarded and (with the crop glue) cropped
    if (!img_dev && g >= 2) OCRB_CUDA(cudaStreamWaitEvent(s_copy, ws->img_free[g & 1], 0));
    for (int c0 = 0, bc = 0; c0 < gn && rc == OCRB_OK; c0 += bc) {
      bc = pipeline_next_chunk(gn - c0, chunk, !img_dev, &ramp);
      const uint8_t *src = images + (size_t)(g0 + c0) * HW;
      if (!img_dev) {
        uint8_t *dst = ws->images[g & 1].as<uint8_t>() + (size_t)c0 * HW;
        OCRB_CUDA(cudaMemcpyAsync(dst, src, (size_t)bc * HW, cudaMemcpyHostToDevice, s_copy));
        OCRB_CUDA(cudaEventRecord(ws->copied[g & 1], s_copy));
        OCRB_CUDA(cudaStreamWaitEvent(s_fwd, ws->copied[g & 1], 0));
        src = dst;
      }
      cudaStream_t saved = ctx->stream;
      ctx->stream = s_fwd;
      // experiment knob: leave a few SMs to the post-processing stream while the persistent convolution kernels run
      static const int sm_leave = getenv("OCRB_SM_LEAVE") ? atoi(getenv("OCRB_SM_LEAVE")) : 0;
      const int saved_limit = ctx->sm_limit;
      if (sm_leave > 0 && !serial) ctx->sm_limit = ctx->sm_count - sm_leave;
      if (split) ctx->sm_limit = ws->fwd_sms;
      rc = det_forward_device(det, src, OCRB_U8, bc, H, W, prob + (size_t)c0 * HW, bf16 ? bitmap + (size_t)c0 * HW : nullptr, (float)prm.thresh);
      if (rc == OCRB_OK && !bf16) rc = launch_binarize(ctx, prob + (size_t)c0 * HW, (int64_t)bc * HW, (float)prm.thresh, bitmap + (size_t)c0 * HW);
      ctx->stream = saved;
      ctx->sm_limit = saved_limit;
    }
    if (rc == OCRB_OK) OCRB_CUDA(cudaEventRecord(ws->fwd_done[g & 1], s_fwd));
    if (rc == OCRB_OK && !img_dev && crop_k == 0) OCRB_CUDA(cudaEventRecord(ws->img_free[g & 1], s_fwd));
    return rc;
  };

  ocrb_polygons *res = polygons_new();
  std::vector<int64_t> group_kept(n_groups, 0);
  int rc = enqueue_forward(0);
  auto cuda_ok = [&](cudaError_t err) {
    if (err == cudaSuccess) return true;
    set_error("pipeline stream ordering -> %s", cudaGetErrorString(err));
    rc = OCRB_ERR_CUDA;
    return false;
  };
  for (int g = 0; g < n_groups && rc == OCRB_OK; ++g) {
    if (g + 1 < n_groups) rc = enqueue_forward(g + 1);
    if (rc != OCRB_OK) break;
    const int g0 = g * group, gn = B - g0 < group ? B - g0 : group;
    // with the SM partition, a group that has a forward to run beside goes to the small partition's stream
    const bool small = split && g + 1 < n_groups;
    cudaStream_t s_ppg = small ? ws->pp_small : s_pp;
    cudaError_t e = cudaStreamWaitEvent(s_ppg, ws->fwd_done[g & 1], 0);
    if (e == cudaSuccess && split) e = cudaStreamWaitEvent(s_ppg, g == 0 ? ws->pp_ready : ws->pp_chain, 0);  // adjust copied / workspace free
    if (e != cudaSuccess) { set_error("cudaStreamWaitEvent -> %s", cudaGetErrorString(e)); rc = OCRB_ERR_CUDA; break; }
    ctx->stream = s_ppg;
    ocrb_polygons *part = polygons_new();
    rc = postproc_device(ctx, ws->prob[g & 1].as<float>(), ws->bitmap[g & 1].as<uint8_t>(), ws->adjust.as<double>() + (size_t)g0 * 2, gn, H, W, prm, part);
    const int64_t n_kept = rc == OCRB_OK ? ocrb_polygons_image_offsets(part)[gn] : 0;
    if (rc == OCRB_OK) polygons_append(res, part);
    ocrb_polygons_free(part);
    if (rc == OCRB_OK && crop_k > 0) {
      // crop glue: kept boxes (device, result order) -> 28x28 tiles from the group's source images -> classes;
      // queued behind the post-processing, read on the host after the final synchronisation
      group_kept[g] = n_kept;
      if (n_kept > 0) {
        const int2 *boxes;
        const int *box_image;
        postproc_kept_boxes(ctx, &boxes, &box_image);
        const int64_t n_tiles = n_kept * crop_k;
        if ((rc = ws->crops[g & 1].reserve((size_t)n_tiles * 784)) != OCRB_OK) break;
        if ((rc = ws->crop_cls[g & 1].reserve((size_t)n_tiles * 4)) != OCRB_OK) break;
        if ((rc = ws->cls_host[g].reserve((size_t)n_tiles * 4)) != OCRB_OK) break;
        const uint8_t *src = img_dev ? images + (size_t)g0 * HW : ws->images[g & 1].as<uint8_t>();
        // the tile / class buffers of this parity are free once group g - 2's glyph net has read them
        if (split && g >= 2 && !cuda_ok(cudaStreamWaitEvent(s_ppg, ws->rec_done[g & 1], 0))) break;
        if ((rc = launch_crop_glyphs(ctx, src, H, W, boxes, box_image, (int)n_kept, crop_k, ws->crops[g & 1].as<uint8_t>())) != OCRB_OK) break;
        // the glyph net is tensor-core work: with the partition it follows the next group's forward on the large one
        cudaStream_t s_rec = small ? s_fwd : s_ppg;
        if (small) {
          if (!cuda_ok(cudaEventRecord(ws->crop_done, s_ppg)) || !cuda_ok(cudaStreamWaitEvent(s_rec, ws->crop_done, 0))) break;
          ctx->stream = s_rec;
          ctx->sm_limit = ws->fwd_sms;
        }
        // one glyph net, one set of activations: its runs are ordered across the two streams
        if (split && g >= 1 && !cuda_ok(cudaStreamWaitEvent(s_rec, ws->rec_done[(g - 1) & 1], 0))) break;
        rc = rec_forward_device(rec, ws->crops[g & 1].p, 1, (int)n_tiles, nullptr, ws->crop_cls[g & 1].as<int32_t>(), nullptr);
        ctx->stream = s_ppg;
        ctx->sm_limit = 0;
        if (rc != OCRB_OK) break;
        e = cudaMemcpyAsync(ws->cls_host[g].p, ws->crop_cls[g & 1].p, (size_t)n_tiles * 4, cudaMemcpyDeviceToHost, s_rec);
        if (e != cudaSuccess) { set_error("cudaMemcpyAsync -> %s", cudaGetErrorString(e)); rc = OCRB_ERR_CUDA; break; }
        if (split && !cuda_ok(cudaEventRecord(ws->rec_done[g & 1], s_rec))) break;
      }
      if (!img_dev) {
        e = cudaEventRecord(ws->img_free[g & 1], s_ppg);
        if (e != cudaSuccess) { set_error("cudaEventRecord -> %s", cudaGetErrorString(e)); rc = OCRB_ERR_CUDA; break; }
      }
    }
    if (split && !cuda_ok(cudaEventRecord(ws->pp_chain, s_ppg))) break;
    ctx->stream = s_pp;
  }
  ctx->stream = s_pp;
  ctx->sm_limit = 0;
  if (split) cudaStreamSynchronize(ws->pp_small);
  if (rc == OCRB_OK) rc = sync(ctx);
  cudaStreamSynchronize(ws->fwd);
  cudaStreamSynchronize(ws->copy);
  if (rc == OCRB_OK && bf16) rc = det_check_err(det);
  if (rc != OCRB_OK) {
    cudaStreamSynchronize(ctx->stream);
    ocrb_polygons_free(res);
    return rc;
  }
  if (crop_k > 0) polygons_set_glyph_classes(res, crop_k, ws->cls_host, group_kept);
  *out = res;
  return OCRB_OK;
}

extern "C" int ocrb_detect_and_recognize(ocrb_det *det, ocrb_rec *rec, const uint8_t *images, const double *adjust, int B, int H,
                                         int W, const ocrb_postproc_params *params, const uint8_t *glyphs, int n_glyphs,
                                         int32_t *glyph_argmax, ocrb_polygons **out) {
  return run_pipeline(det, rec, images, adjust, B, H, W, params, glyphs, n_glyphs, glyph_argmax, 0, out);
}

extern "C" int ocrb_detect_and_read(ocrb_det *det, ocrb_rec *rec, const uint8_t *images, const double *adjust, int B, int H, int W,
                                    const ocrb_postproc_params *params, int glyphs_per_polygon, ocrb_polygons **out) {
  OCRB_REQUIRE(glyphs_per_polygon > 0, "glyphs_per_polygon must be positive");
  return run_pipeline(det, rec, images, adjust, B, H, W, params, nullptr, 0, nullptr, glyphs_per_polygon, out);
}
