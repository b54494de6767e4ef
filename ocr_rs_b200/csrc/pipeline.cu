// run_text_detection's device part for a batch (text_detection/mod.rs:46-67; the batched form
// is get_model_accuracy, :188-204): u8 images -> detector -> binarize -> post-processing,
// plus glyph recognition (char_recognition/mod.rs:39-68) of a caller-provided crop set in the
// same call (the reference has no polygon -> crop glue).
//
// Images are processed in chunks so the activation workspace stays bounded; each chunk is
// H2D copy -> forward (the BF16 head writes the probability map AND the bitmap) -> post-proc.
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
  DevBuf images, prob, bitmap, adjust, glyphs, argmax;
};
static PipelineWorkspace *g_ws[16] = {nullptr};

static PipelineWorkspace *get_ws(ocrb_ctx *ctx) {
  PipelineWorkspace *&w = g_ws[ctx->device & 15];
  if (!w) w = new PipelineWorkspace();
  return w;
}

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
  PipelineWorkspace *ws = get_ws(ctx);
  const int64_t HW = (int64_t)H * W;
  const bool bf16 = det_mode(det) == OCRB_MODE_BF16;
  // chunk: <= 32 images and <= 2^31 pixels for the post-processing index arithmetic
  int chunk = bf16 ? 32 : 4;
  while ((int64_t)chunk * HW >= ((int64_t)1 << 31) && chunk > 1) chunk /= 2;
  if (chunk > B) chunk = B;
  const bool img_dev = is_device_ptr(images);
  OCRB_TRY(ws->prob.reserve((size_t)chunk * HW * 4));
  OCRB_TRY(ws->bitmap.reserve((size_t)chunk * HW));
  OCRB_TRY(ws->adjust.reserve((size_t)B * 16));
  if (!img_dev) OCRB_TRY(ws->images.reserve((size_t)chunk * HW));
  OCRB_CUDA(cudaMemcpyAsync(ws->adjust.p, adjust, (size_t)B * 16, cudaMemcpyDefault, ctx->stream));

  // glyph recognition first: its kernels queue behind nothing and overlap the first H2D copy
  if (n_glyphs > 0) {
    const void *g = glyphs;
    if (!is_device_ptr(glyphs)) {
      OCRB_TRY(ws->glyphs.reserve((size_t)n_glyphs * 784));
      OCRB_CUDA(cudaMemcpyAsync(ws->glyphs.p, glyphs, (size_t)n_glyphs * 784, cudaMemcpyHostToDevice, ctx->stream));
      g = ws->glyphs.p;
    }
    int32_t *am = glyph_argmax;
    if (glyph_argmax && !is_device_ptr(glyph_argmax)) {
      OCRB_TRY(ws->argmax.reserve((size_t)n_glyphs * 4));
      am = ws->argmax.as<int32_t>();
    }
    OCRB_TRY(rec_forward_device(rec, g, 1, n_glyphs, nullptr, am, nullptr));
    if (glyph_argmax && am != glyph_argmax)
      OCRB_CUDA(cudaMemcpyAsync(glyph_argmax, am, (size_t)n_glyphs * 4, cudaMemcpyDeviceToHost, ctx->stream));
  }

  ocrb_polygons *res = polygons_new();
  int rc = OCRB_OK;
  for (int b0 = 0; b0 < B && rc == OCRB_OK; b0 += chunk) {
    const int bc = B - b0 < chunk ? B - b0 : chunk;
    const uint8_t *src = images + (size_t)b0 * HW;
    if (!img_dev) {
      cudaError_t e = cudaMemcpyAsync(ws->images.p, src, (size_t)bc * HW, cudaMemcpyHostToDevice, ctx->stream);
      if (e != cudaSuccess) { set_error("H2D image copy -> %s", cudaGetErrorString(e)); rc = OCRB_ERR_CUDA; break; }
      src = ws->images.as<uint8_t>();
    }
    if (bf16) {
      rc = det_forward_device(det, src, OCRB_U8, bc, H, W, ws->prob.as<float>(), ws->bitmap.as<uint8_t>(), (float)prm.thresh);
    } else {
      rc = det_forward_device(det, src, OCRB_U8, bc, H, W, ws->prob.as<float>(), nullptr, (float)prm.thresh);
      if (rc == OCRB_OK) rc = launch_binarize(ctx, ws->prob.as<float>(), (int64_t)bc * HW, (float)prm.thresh, ws->bitmap.as<uint8_t>());
    }
    if (rc != OCRB_OK) break;
    ocrb_polygons *part = polygons_new();
    rc = postproc_device(ctx, ws->prob.as<float>(), ws->bitmap.as<uint8_t>(), ws->adjust.as<double>() + (size_t)b0 * 2, bc, H, W, prm, part);
    if (rc == OCRB_OK) polygons_append(res, part);
    ocrb_polygons_free(part);
  }
  if (rc == OCRB_OK) rc = sync(ctx);
  if (rc == OCRB_OK && bf16) rc = det_check_err(det);
  if (rc != OCRB_OK) {
    cudaStreamSynchronize(ctx->stream);
    ocrb_polygons_free(res);
    return rc;
  }
  *out = res;
  return OCRB_OK;
}
