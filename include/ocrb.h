/*
 * ocrb.h — C ABI of the B200-native text-detection / glyph-recognition hot path.
 *
 * This is the drop-in boundary for lazareviczoran/ocr-rs: every entry point replaces one
 * reference function (cited as file:line into the reference tree) and is what a Rust
 * `ocrb-sys` crate (cc/bindgen) would bind — see INTEGRATION.md for the Rust side.
 *
 * Conventions
 *   - plain C, no C++/torch types; every call returns 0 (OCRB_OK) or a negative error code;
 *     ocrb_last_error() gives the thread-local message.  Nothing throws across the ABI.
 *   - one ocrb_ctx per (device, stream); one ctx per host thread; no hidden global state.
 *   - every data pointer may be HOST memory (pageable or pinned) or DEVICE memory on the
 *     ctx's device; the library inspects the pointer (cudaPointerGetAttributes) and stages
 *     host buffers itself.
 *   - stream ordering: all work is launched on the ctx's own non-blocking stream(s) and every
 *     entry point returns only after that work has completed, so OUTPUTS (host or device) are
 *     complete on return.  A device INPUT must be complete before the call — synchronise its
 *     producer, or call ocrb_ctx_wait_stream(ctx, producer_stream) first, which orders everything
 *     queued on that stream so far before the ctx's later work without blocking the host.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     OCRB_ERR_CUDA.
 *   - layouts follow the reference: images / maps are row-major [B][H][W]; points are
 *     (x, y) pairs; weights are OIHW float32 under the reference's VarStore names
 *     (SURVEY.md Appendix B).
 */
#ifndef OCRB_H
#define OCRB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCRB_VERSION 100

enum {
  OCRB_OK = 0,
  OCRB_ERR_INVALID = -1,   /* bad argument / shape / missing weight name */
  OCRB_ERR_CUDA = -2,      /* CUDA runtime or driver error, or no device */
  OCRB_ERR_CAPACITY = -3,  /* caller-provided output buffer too small */
  OCRB_ERR_INTERNAL = -4
};

/* arithmetic mode of the detector (north_star tolerances: FP32 1e-4, BF16 1e-2) */
enum { OCRB_MODE_FP32 = 0, OCRB_MODE_BF16 = 1 };
/* element type of detector input images */
enum { OCRB_U8 = 0, OCRB_F32 = 1 };

typedef struct ocrb_ctx ocrb_ctx;
typedef struct ocrb_det ocrb_det;
typedef struct ocrb_rec ocrb_rec;
typedef struct ocrb_polygons ocrb_polygons;
typedef struct ocrb_varstore ocrb_varstore;

/* ---- context ------------------------------------------------------------------------
 * replaces the process-global `DEVICE` (main.rs:26-28) with an explicit handle */
int ocrb_version(void);
const char *ocrb_last_error(void);
int ocrb_device_count(int *count);
int ocrb_ctx_create(int device, ocrb_ctx **out);
int ocrb_ctx_destroy(ocrb_ctx *ctx);
int ocrb_ctx_synchronize(ocrb_ctx *ctx);
void *ocrb_ctx_stream(ocrb_ctx *ctx); /* cudaStream_t the ctx launches on */
/* orders the work queued so far on `producer_stream` (a cudaStream_t; NULL = the legacy default
 * stream) before everything this ctx launches afterwards.  Does not block the host. */
int ocrb_ctx_wait_stream(ocrb_ctx *ctx, void *producer_stream);
int ocrb_ctx_device(ocrb_ctx *ctx);
/* number of kernels this ctx has launched since creation (bench.py "gpu_launches") */
int64_t ocrb_ctx_launch_count(ocrb_ctx *ctx);
/* per-launch CUDA-event timeline (replaces the measure_time! macro, macros.rs:46-71).
 * begin: synchronise and start recording one event per launch/copy on the ctx stream.
 * end: synchronise, stop, and write "name count total_ms\n" lines (one per kernel name, in
 * first-seen order) into buf; *needed receives the byte count including the NUL — call with
 * buf == NULL to size the buffer (the timeline is kept until the next begin). */
int ocrb_ctx_profile_begin(ocrb_ctx *ctx);
int ocrb_ctx_profile_end(ocrb_ctx *ctx, char *buf, size_t cap, size_t *needed);

/* ---- image_ops ----------------------------------------------------------------------
 * image_ops::preprocess_image (image_ops.rs:188-220) minus the file decode:
 * RGBA8 [src_h][src_w][4] -> aspect-preserving Triangle resize (image 0.23.11) -> Rec.709
 * luma (truncating) -> zero-padded top-left into [H][W] u8; adjust = resized / original. */
int ocrb_resize_dims(int src_w, int src_h, int W, int H, int *resized_w, int *resized_h);
int ocrb_preprocess_rgba(ocrb_ctx *ctx, const uint8_t *rgba, int src_w, int src_h, int W, int H,
                         uint8_t *out_gray, double *adjust_x, double *adjust_y);
/* the same for a batch in ONE fused launch (the reference's u8 intermediate image lives in shared memory only):
 * image i = RGBA8 [src_h[i]][src_w[i]][4] at byte offset src_offsets[i] (a multiple of 4) of `rgba`;
 * out_gray [n][H][W], adjust [n][2] = (adjust_x, adjust_y) */
int ocrb_preprocess_rgba_batch(ocrb_ctx *ctx, const uint8_t *rgba, const int64_t *src_offsets, const int *src_w, const int *src_h, int n,
                               int W, int H, uint8_t *out_gray, double *adjust);
/* ---- file decode: `image::open(file)?.into_rgba()` / `.into_luma()` (image_ops.rs:193, :78; image 0.23.11 ->
 * jpeg-decoder 0.1.20 / png 0.16.7) ----------------------------------------------------------------------
 * JPEG (baseline + progressive Huffman, 8 bit, grey or YCbCr with sampling ratios 1 and 2) and PNG (8-bit and
 * sub-byte grey / palette / RGB / RGBA, tRNS, non-interlaced) are decoded; anything else returns OCRB_ERR_INVALID
 * (the reference returns Err from image::open).  The serial entropy decode runs on host threads (one image
 * each), the inverse DCT, chroma upsampling and colour conversion run on the device for the whole batch; results
 * are bit-identical to the reference's decoder (pinned on its preprocessed_img*.png fixtures).
 * ocrb_image_info needs no device. */
enum { OCRB_PIXELS_RGBA = 0, OCRB_PIXELS_LUMA = 1 };
int ocrb_image_info(const uint8_t *file, size_t size, int *width, int *height, int *channels);
/* host-only test hook: what the host stage hands to the device for one file (JPEG: int16 DCT coefficients,
 * components concatenated, natural order; PNG: pixels [h][w][channels]); out == NULL sizes the buffer */
int ocrb_debug_decode_host(const uint8_t *file, size_t size, void *out, size_t cap, size_t *needed);
/* files[i] (sizes[i] bytes) -> pixels at byte offset out_offsets[i] of `out` (host or device):
 * RGBA8 [h][w][4] (offsets multiples of 4) or luma [h][w] */
int ocrb_decode_images(ocrb_ctx *ctx, const uint8_t *const *files, const size_t *sizes, int n, int format,
                       const int64_t *out_offsets, uint8_t *out);
/* image_ops::preprocess_image(file, (W, H)) (image_ops.rs:188-220) for a batch of encoded files: decode ->
 * resize -> luma -> pad with the decoded pixels kept in HBM; out_gray [n][H][W], adjust [n][2] */
int ocrb_preprocess_files(ocrb_ctx *ctx, const uint8_t *const *files, const size_t *sizes, int n, int W, int H,
                          uint8_t *out_gray, double *adjust);
/* image_ops::convert_image_to_tensor + to_kind(Float) (image_ops.rs:350-364,
 * text_detection/mod.rs:46-49): u8 -> f32, no scaling. */
int ocrb_convert_image_to_tensor(ocrb_ctx *ctx, const uint8_t *image, int64_t n, float *out);
/* image_ops::convert_tensor_to_image (image_ops.rs:367-381): f32 -> u8 by truncation
 * (to_kind(Uint8)); `scale` is applied first in f32 (mod.rs:57 uses 255). */
int ocrb_convert_tensor_to_image(ocrb_ctx *ctx, const float *tensor, int64_t n, float scale, uint8_t *out);
/* image_ops::load_image_as_tensor (image_ops.rs:73-85) minus the decode: u8 -> f32 / 255 */
int ocrb_load_image_as_tensor(ocrb_ctx *ctx, const uint8_t *luma, int64_t n, float *out);

/* ---- text_detection::model ----------------------------------------------------------
 * resnet18(&vs.root()) + vs.load(file) (model.rs:65-156, text_detection/mod.rs:35-44):
 * takes the 121 named OIHW float32 tensors; batch-norm is folded
 * internally.  numel[i] is checked against the shape the name implies. */
int ocrb_det_create(ocrb_ctx *ctx, int n_tensors, const char *const *names,
                    const float *const *data, const int64_t *numel, int mode, ocrb_det **out);
int ocrb_det_destroy(ocrb_det *det);
/* net.forward_t(images.view(B,1,H,W), false) (text_detection/mod.rs:52-54, :196-197).
 * images: [B][H][W] u8 or f32 raw grey levels 0..255 (no normalisation).
 * H and W must be multiples of 32 (model.rs:126-137).  prob: [B][H][W] f32. */
int ocrb_det_forward(ocrb_det *det, const void *images, int dtype, int B, int H, int W, float *prob);
/* test hook: copies an intermediate feature map (NCHW f32) out of the last forward.
 * name in {"stem","x1","x2","x3","x4","fuse","bin1"}; numel must match. */
int ocrb_det_tap(ocrb_det *det, const char *name, float *out, int64_t numel);

/* ---- text_detection::metrics --------------------------------------------------------
 * binarize (metrics.rs:129-131): out = pred > (float)thresh, u8 {0,1}. */
int ocrb_binarize(ocrb_ctx *ctx, const float *pred, int64_t n, double thresh, uint8_t *out);
/* box_score_fast (metrics.rs:150-184): pred is [dim_m2][dim_m1] f32; xy = n_pts (x,y) u32-range
 * points.  The reference's swapped w/h clamps are kept literally. */
int ocrb_box_score_fast(ocrb_ctx *ctx, const float *pred, int dim_m2, int dim_m1,
                        const int32_t *xy, int n_pts, double *score);
/* get_min_area_bounding_box (metrics.rs:133-148): box_xy receives 4 (x,y) corners */
int ocrb_min_area_bounding_box(ocrb_ctx *ctx, const int32_t *xy, int n_pts, int32_t *box_xy, double *sside);
/* polygon::expand_polygon (polygon.rs:51-56). *n_out = 0 means None. */
int ocrb_expand_polygon(ocrb_ctx *ctx, const int32_t *xy, int n_pts, double factor,
                        int32_t *out_xy, int out_cap_pts, int *n_out);
/* polygon::clip_polygon (polygon.rs:13-42) / shrink_polygon (polygon.rs:44-49) on the HOST, no device needed:
 * offset_type shrink != 0 -> negative distance.  The reference uses it only to prepare training targets
 * (image_ops.rs:222-277), on the CPU; it is host code here as well — compiled from the same offset / union functions
 * the device unclip runs, so the CPU tests hold that source to the reference's gt_shrinked fixtures.  *n_out = 0 means
 * None; *distance (optional) = the signed offset distance.  out_cap_pts >= 6 * n_pts + 32 always suffices. */
int ocrb_clip_polygon(const int32_t *xy, int n_pts, double factor, int shrink,
                      int32_t *out_xy, int out_cap_pts, int *n_out, double *distance);

typedef struct ocrb_postproc_params {
  double thresh;        /* 0.6  metrics.rs:38  */
  double box_thresh;    /* 0.7  metrics.rs:64  */
  double min_size;      /* 5.0  metrics.rs:66  */
  double unclip_factor; /* 2.0  metrics.rs:103 */
} ocrb_postproc_params;
void ocrb_postproc_default_params(ocrb_postproc_params *p);

/* get_boxes_and_box_scores (metrics.rs:37-56): pred [B][H][W] f32 (the reference's
 * [B,1,H,W]), adjust [B][2] f64.  params == NULL selects the reference constants.
 * The result handle owns host memory; free with ocrb_polygons_free. */
int ocrb_get_boxes_and_box_scores(ocrb_ctx *ctx, const float *pred, const double *adjust,
                                  int B, int H, int W, const ocrb_postproc_params *params,
                                  ocrb_polygons **out);
/* get_polygons_from_bitmap (metrics.rs:58-127): one image, caller-provided bitmap */
int ocrb_get_polygons_from_bitmap(ocrb_ctx *ctx, const float *pred, const uint8_t *bitmap,
                                  const double *adjust, int H, int W,
                                  const ocrb_postproc_params *params, ocrb_polygons **out);
/* PolygonScores accessors (metrics.rs:32-35).  Polygons of image b are
 * [image_offsets[b], image_offsets[b+1]); polygon p owns points
 * [point_offsets[p], point_offsets[p+1]) of xy (u32 x,y pairs); scores[p] is f64. */
int ocrb_polygons_num_images(const ocrb_polygons *p);
const int64_t *ocrb_polygons_image_offsets(const ocrb_polygons *p);
const int64_t *ocrb_polygons_point_offsets(const ocrb_polygons *p);
const uint32_t *ocrb_polygons_xy(const ocrb_polygons *p);
const double *ocrb_polygons_scores(const ocrb_polygons *p);
/* per image: contours, >=4 DP points, >= box_thresh, kept, dropped (empty offset, D11) */
const int64_t *ocrb_polygons_stats(const ocrb_polygons *p);
void ocrb_polygons_free(ocrb_polygons *p);

/* ---- evaluation metrics (metrics.rs:191-394; host code, no device needed) -----------------------------
 * MetricsItem (metrics.rs:22-30) */
typedef struct ocrb_metrics_item {
  double precision, recall, hmean;
  int64_t gt_care, det_care, det_matched;
} ocrb_metrics_item;
/* get_intersection / get_union / get_intersection_over_union (metrics.rs:375-389): polygons as u32 (x,y) rings
 * without the closing point; either output may be NULL */
int ocrb_polygon_iou(const uint32_t *a_xy, int n_a, const uint32_t *b_xy, int n_b, double *intersection, double *iou);
/* evaluate_image (metrics.rs:251-372): polygon g owns points [gt_offsets[g], gt_offsets[g+1]) of gt_xy (likewise the
 * detections); ignore_flags[g] != 0 marks a don't-care ground-truth polygon.  validate_measure (metrics.rs:191-218) is
 * this call per image on the detections with score >= 0.6. */
int ocrb_evaluate_image(const int64_t *gt_offsets, const uint32_t *gt_xy, int n_gt, const uint8_t *ignore_flags,
                        const int64_t *det_offsets, const uint32_t *det_xy, int n_det, ocrb_metrics_item *out);
/* combine_results / gather_measure (metrics.rs:220-249) -> (precision, recall, hmean) */
int ocrb_combine_results(const ocrb_metrics_item *items, int n, double *precision, double *recall, double *hmean);

/* fine-grained test hooks of the contour stage (imageproc find_contours at
 * metrics.rs:78-81 and approximate_polygon_dp at :87-95) */
/* 8-connected foreground labels, canonical raster-order numbering 1..n, 0 = background */
int ocrb_ccl_labels(ocrb_ctx *ctx, const uint8_t *bitmap, int B, int H, int W, int32_t *labels, int32_t *n_components);
/* all border chains of one bitmap, in the reference's order.  chain c owns points
 * [offsets[c], offsets[c+1]) of xy; types[c] = 0 outer / 1 hole.  Pass NULL outputs with
 * zero capacities to query sizes through n_contours / n_points. */
int ocrb_find_contours(ocrb_ctx *ctx, const uint8_t *bitmap, int H, int W,
                       int64_t *offsets, uint8_t *types, int64_t contour_cap,
                       int32_t *xy, int64_t point_cap, int64_t *n_contours, int64_t *n_points);
/* eps = 0.01 * arc_length (0 -> 0.01), DP, duplicated last point dropped (metrics.rs:87-95) */
int ocrb_approx_polygon(ocrb_ctx *ctx, const int32_t *chain_xy, int64_t n_pts,
                        int32_t *out_xy, int64_t out_cap_pts, int64_t *n_out);

/* host-only test hook (no device needed): the tiling the 3x3 convolution kernel (csrc/conv_halo.cu) would choose for an
 * Ho x Wo map.  mode 0 = N tile 128 (linear sub-tiles), 1 = N tile 64 with the TMA-store epilogue, 2 = pair mode.
 * out[10] = {PW, TH, TW, sub_rows, sub_stride, a_stage_bytes, a_stages, b_stages, obufs * obuf_bytes, dynamic smem bytes} */
int ocrb_debug_conv_geometry(int Ho, int Wo, int mode, int *out);

/* host-only test hook (no device needed): the batching plan ocrb_detect_and_read / ocrb_detect_and_recognize would follow
 * for B images of H x W (the batched loop of text_detection/mod.rs:188-204): images per post-processing group, and the
 * forward chunks in order (bf16: mode of the detector; host_images: the images come from host memory, so the first
 * chunks ramp up 16, 48, 144 to hide all but the first copy).  chunks[cap]; OCRB_ERR_CAPACITY if the plan has more. */
int ocrb_debug_pipeline_plan(int B, int H, int W, int bf16, int host_images, int *group, int *chunks, int cap, int *n_chunks);

/* host-only test hook (no device needed): get_min_area_bounding_box (metrics.rs:133-148) computed on the host by the
 * function the device unclip kernel runs (compiled __host__ __device__) — box_xy receives 4 (x,y) corners.  The product
 * entry point is ocrb_min_area_bounding_box. */
int ocrb_debug_min_area_bounding_box_host(const int32_t *xy, int n_pts, int32_t *box_xy, double *sside);

/* host-only test hook (no device needed): approximate_polygon_dp with eps = 1 % of the closed arc length, as
 * metrics.rs:87-95 uses it (the effective result: open-chain Douglas-Peucker minus its last point), run on the host by the
 * device kernel's own statements.  chain_xy: n_pts border pixels (x, y in 0..65535).  The product entry point is
 * ocrb_approx_polygon. */
int ocrb_debug_approx_polygon_host(const int32_t *chain_xy, int64_t n_pts, int32_t *out_xy, int64_t out_cap_pts, int64_t *n_out);

/* ---- char_recognition ---------------------------------------------------------------
 * Net::new + vs.load (char_recognition/model.rs:12-25, mod.rs:43-45).  Names: canonical
 * "conv1.weight" ... "fc2.bias" or the de-duplicated VarStore names. */
int ocrb_rec_create(ocrb_ctx *ctx, int n_tensors, const char *const *names,
                    const float *const *data, const int64_t *numel, ocrb_rec **out);
int ocrb_rec_destroy(ocrb_rec *rec);
/* Net::forward_t(xs, false) + softmax(-1, Double) + topk(1) (model.rs:27-39, mod.rs:53-56,
 * utils.rs:28-43).  glyphs: [B][784] f32 in [0,1].  Any of logits [B][62] f32,
 * argmax [B] i32, prob [B] f64 may be NULL. */
int ocrb_rec_forward(ocrb_rec *rec, const float *glyphs, int B, float *logits, int32_t *argmax, double *prob);
/* same on raw u8 glyphs (load_image_as_tensor's /255 fused in) */
int ocrb_rec_forward_u8(ocrb_rec *rec, const uint8_t *glyphs, int B, float *logits, int32_t *argmax, double *prob);
/* utils::VALUES (utils.rs:7): class index -> character */
char ocrb_class_to_char(int cls);

/* ---- model files ---------------------------------------------------------------------
 * vs.load(file) (text_detection/mod.rs:40-44, char_recognition/mod.rs:43-45): native reader of
 * the libtorch archive tch's VarStore::save writes (utils.rs:55-63) — ZIP of STORED entries,
 * `data.pkl` + raw storages.  Host-only; tensors are returned as float32 under their VarStore
 * names (f64 / f16 / bf16 / integer storages are converted). */
int ocrb_varstore_open(const char *path, ocrb_varstore **out);
int ocrb_varstore_count(const ocrb_varstore *vs);
const char *ocrb_varstore_name(const ocrb_varstore *vs, int i);
int ocrb_varstore_tensor(const ocrb_varstore *vs, int i, const float **data, int64_t *numel, const int64_t **shape, int *ndim);
void ocrb_varstore_close(ocrb_varstore *vs);
/* resnet18(&vs.root()) + vs.load(path) / Net::new(&vs.root()) + vs.load(path) in one call */
int ocrb_det_create_from_file(ocrb_ctx *ctx, const char *path, int mode, ocrb_det **out);
int ocrb_rec_create_from_file(ocrb_ctx *ctx, const char *path, ocrb_rec **out);

/* ---- pipeline -----------------------------------------------------------------------
 * run_text_detection's device part for a batch (text_detection/mod.rs:46-67, :188-204):
 * images u8 [B][H][W] -> detector -> post-processing; plus glyph recognition of
 * `n_glyphs` 28x28 u8 crops in the same call (the reference has no crop glue).
 * Host pointers are staged through pinned memory; result as above; argmax may be NULL. */
int ocrb_detect_and_recognize(ocrb_det *det, ocrb_rec *rec, const uint8_t *images, const double *adjust,
                              int B, int H, int W, const ocrb_postproc_params *params,
                              const uint8_t *glyphs, int n_glyphs, int32_t *glyph_argmax,
                              ocrb_polygons **out);

/* detection + the polygon -> glyph crop glue + recognition in one call.  The reference stops at the polygons
 * (character segmentation is an open item of its README.md:20-26) and feeds its recognition net ready-made
 * 28x28 files (image_ops.rs:73-85); this entry point closes the gap on the device: every kept polygon's min-area
 * rectangle (metrics.rs:133-148) is cut out of its source image into `glyphs_per_polygon` cells along the
 * rectangle's longer side, each cell is resized to 28x28 with preprocess_image's Triangle filter (crop spec v1,
 * oracle/postproc_oracle.c orc_crop_glyphs) and classified.  Result: the polygons as above plus
 * glyph_classes [n_polygons][glyphs_per_polygon] (class index -> character: ocrb_class_to_char). */
int ocrb_detect_and_read(ocrb_det *det, ocrb_rec *rec, const uint8_t *images, const double *adjust, int B, int H, int W,
                         const ocrb_postproc_params *params, int glyphs_per_polygon, ocrb_polygons **out);
int ocrb_polygons_glyphs_per_polygon(const ocrb_polygons *p);
const int32_t *ocrb_polygons_glyph_classes(const ocrb_polygons *p);
/* test hook of the crop stage: image u8 [H][W], boxes [n][4] (x,y) corners TL,TR,BR,BL as ocrb_min_area_bounding_box
 * returns them -> out_glyphs u8 [n * glyphs_per_box][784] */
int ocrb_crop_glyphs(ocrb_ctx *ctx, const uint8_t *image, int H, int W, const int32_t *boxes_xy, int n_boxes, int glyphs_per_box,
                     uint8_t *out_glyphs);

/* ---- several GPUs of one box, one process -----------------------------------------------------
 * The batch contract of the reference's evaluation loop (text_detection/mod.rs:188-204: images [n][H][W] +
 * adjust [n][2] -> one PolygonScores) across devices: contiguous index shards (the first n % G shards hold one
 * image more), one host thread per device inside the library, each with its own ctx + detector + recognition
 * net (weights replicated), results appended into ONE host CSR in image order.  No collective.
 * All buffers are HOST memory; allocate them with ocrb_host_alloc (page-locked) for full copy / compute overlap. */
typedef struct ocrb_shards ocrb_shards;
int ocrb_shards_create(const int *devices, int n_devices,
                       int n_det, const char *const *det_names, const float *const *det_data, const int64_t *det_numel, int mode,
                       int n_rec, const char *const *rec_names, const float *const *rec_data, const int64_t *rec_numel, /* n_rec = 0: no recognition net */
                       ocrb_shards **out);
int ocrb_shards_create_from_files(const int *devices, int n_devices, const char *det_path, const char *rec_path /* may be NULL */,
                                  int mode, ocrb_shards **out);
int ocrb_shards_destroy(ocrb_shards *s);
int ocrb_shards_count(const ocrb_shards *s);
int ocrb_shards_device(const ocrb_shards *s, int i);
int64_t ocrb_shards_launch_count(const ocrb_shards *s); /* kernels launched by all shards since creation */
/* [first, first + count) of shard `shard` out of `n_shards` */
int ocrb_shard_range(int64_t n_items, int shard, int n_shards, int64_t *first, int64_t *count);
/* ocrb_detect_and_recognize over all devices of `s`; glyphs are split the same way as the images */
int ocrb_detect_and_recognize_sharded(ocrb_shards *s, const uint8_t *images, const double *adjust, int B, int H, int W,
                                      const ocrb_postproc_params *params, const uint8_t *glyphs, int n_glyphs,
                                      int32_t *glyph_argmax, ocrb_polygons **out);
/* ocrb_detect_and_read over all devices of `s` */
int ocrb_detect_and_read_sharded(ocrb_shards *s, const uint8_t *images, const double *adjust, int B, int H, int W,
                                 const ocrb_postproc_params *params, int glyphs_per_polygon, ocrb_polygons **out);
int ocrb_host_alloc(size_t bytes, void **out);
int ocrb_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* OCRB_H */
