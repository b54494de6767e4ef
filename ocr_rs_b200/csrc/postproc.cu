// Host orchestration of the detection post-processing on one device stream:
//   get_boxes_and_box_scores / get_polygons_from_bitmap  (metrics.rs:37-127)
// binarize -> CCL -> border starts -> chains -> Douglas–Peucker -> (>= 4 points) -> box score
// -> (>= box_thresh) -> unclip -> min-area-rect (>= min_size) -> rescale -> polygons,
// everything on the GPU; the host only reads back element counts between stages to size
// the next stage's arenas, and the final polygon list.
#include "common.cuh"
#include "scan.cuh"

#include <algorithm>

namespace ocrb {

// kernels / launchers defined in the other translation units
int launch_binarize(ocrb_ctx *, const float *, int64_t, float, uint8_t *);
int launch_ccl(ocrb_ctx *, const uint8_t *, int, int, int, int *, bool);
int launch_contour_starts(ocrb_ctx *, const uint8_t *, const int *, int, int, int, uint8_t *, uint8_t *, int4 *, uint8_t *, int *, int *);
int launch_contour_records(ocrb_ctx *, const uint8_t *, const int *, int64_t, int64_t *, uint8_t *);
int launch_contour_count(ocrb_ctx *, int64_t, int *);
int64_t contour_count_slot(int64_t);
int launch_trace_count(ocrb_ctx *, const uint8_t *, int, int, const int64_t *, const uint8_t *, int64_t, int *);
int launch_trace_store(ocrb_ctx *, const uint8_t *, int, int, const int64_t *, const uint8_t *, int64_t, const int64_t *, ushort2 *);
int launch_approx_dp(ocrb_ctx *, const ushort2 *, const int64_t *, int64_t, int *, ushort2 *, int *);
int launch_box_score(ocrb_ctx *, const float *, int, int, int64_t, const int *, const int64_t *, const int64_t *,
                     const ushort2 *, const int *, int, double *, int *);
int launch_unclip_slab_sizes(ocrb_ctx *, const int *, const int *, int, int64_t *);
int clip_polygon_host(const int32_t *, int, double, int, int32_t *, int, int *, double *);
int min_area_bounding_box_host(const int32_t *, int, int32_t *, double *);
int approx_polygon_host(const int32_t *, int64_t, int32_t *, int64_t, int64_t *);
int launch_unclip(ocrb_ctx *, const int *, const int64_t *, const ushort2 *, const int *, int, const double *, double,
                  double, double, const int64_t *, int2 *, int *, uint8_t *, double *, int2 *);
int launch_emit_polygons(ocrb_ctx *, const int *, const int *, const int64_t *, int64_t, int, const uint8_t *, const int *,
                         const int64_t *, const int64_t *, const int2 *, const int *, const double *, const double *,
                         uint32_t *, double *, int64_t *, int *, const int2 *, int2 *);
int launch_flag_ge4(ocrb_ctx *, const int *, int64_t, uint8_t *);
int launch_compact_index(ocrb_ctx *, const uint8_t *, const int *, int64_t, int *);
int launch_kept_sizes(ocrb_ctx *, const uint8_t *, const int *, int, uint8_t *, int *);
int launch_stats(ocrb_ctx *, const int64_t *, const int *, int64_t, int64_t, const int *, int, const uint8_t *, unsigned long long *);
int launch_minrect_hook(ocrb_ctx *, const int2 *, int, double2 *, double2 *, int2 *, double *);
__host__ __device__ inline int unclip_cap_h(int n) { return 6 * n + 32; }

struct PostprocWorkspace {
  DevBuf bitmap, labels, bg_open, hole_traced, flags, bbox, offs, scan_scratch;
  DevBuf start_idx, kind, lengths, chain_off, dp_count, cand_flag, cand_rank;
  DevBuf chain, dp_out, stack;
  DevBuf cand_contour, scores, slab_units, slab_off, out_count, status, kept_flag, kept_pts, kept_rank, pt_off, slabs;
  DevBuf out_xy, out_scores, out_pt_off, out_image, stats, err, adjust;
  DevBuf cand_box, out_box;  // min-area boxes of all candidates / of the kept polygons (TL, TR, BR, BL; map coordinates)
  void release() {
    DevBuf *all[] = {&bitmap, &labels, &bg_open, &hole_traced, &flags, &bbox, &offs, &scan_scratch, &start_idx, &kind,
                     &lengths, &chain_off, &dp_count, &cand_flag, &cand_rank, &chain, &dp_out, &stack, &cand_contour,
                     &scores, &slab_units, &slab_off, &out_count, &status, &kept_flag, &kept_pts, &kept_rank, &pt_off,
                     &slabs, &out_xy, &out_scores, &out_pt_off, &out_image, &stats, &err, &adjust, &cand_box, &out_box};
    for (DevBuf *b : all) b->release();
  }
};

PostprocWorkspace *get_pp(ocrb_ctx *ctx) {
  if (!ctx->pp) ctx->pp = new PostprocWorkspace();
  return ctx->pp;
}
void free_pp(ocrb_ctx *ctx) {
  if (ctx->pp) {
    ctx->pp->release();
    delete ctx->pp;
    ctx->pp = nullptr;
  }
}

template <class T>
static int read_scalar(ocrb_ctx *ctx, const T *dev, T *host) {
  OCRB_CUDA(cudaMemcpyAsync(host, dev, sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}

}  // namespace ocrb

struct ocrb_polygons {
  int n_images = 0;
  std::vector<int64_t> image_offsets;  // n_images + 1
  std::vector<int64_t> point_offsets;  // n_polys + 1
  std::vector<uint32_t> xy;            // 2 * n_points
  std::vector<double> scores;          // n_polys
  std::vector<int64_t> stats;          // 5 * n_images
  int glyphs_per_polygon = 0;          // > 0: filled by ocrb_detect_and_read
  std::vector<int32_t> glyph_classes;  // n_polys * glyphs_per_polygon
  // contour-stage outputs kept for the test hooks
  std::vector<int64_t> chain_offsets;
  std::vector<uint8_t> chain_types;
  std::vector<int32_t> chain_xy;
};

namespace ocrb {

struct ContourStage {
  int64_t n_contours = 0, n_points = 0;
};

// bitmap (device, [B][H][W]) -> contour records + chains + DP polygons in the workspace
static int run_contour_stage(ocrb_ctx *ctx, PostprocWorkspace *ws, const uint8_t *bitmap, int B, int H, int W,
                             ContourStage *st, bool with_dp) {
  const int64_t n = (int64_t)B * H * W;
  OCRB_REQUIRE(W <= 65535 && H <= 65535, "map side must be <= 65535 (got %dx%d)", W, H);
  OCRB_REQUIRE(n < (int64_t)1 << 31, "B*H*W must be < 2^31 (got %lld)", (long long)n);
  OCRB_TRY(ws->labels.reserve(n * 4));
  OCRB_TRY(ws->bg_open.reserve(n));
  OCRB_TRY(ws->hole_traced.reserve(n));
  OCRB_TRY(ws->flags.reserve(n));
  OCRB_TRY(ws->bbox.reserve((size_t)B * H * sizeof(int4) + 16));
  OCRB_TRY(ws->offs.reserve(scan_scratch_elems(n) * 2 * 4 + 64));
  OCRB_TRY(ws->scan_scratch.reserve(scan_scratch_elems(n) * 8));
  OCRB_TRY(launch_ccl(ctx, bitmap, B, H, W, ws->labels.as<int>(), false));
  OCRB_TRY(launch_contour_starts(ctx, bitmap, ws->labels.as<int>(), B, H, W, ws->bg_open.as<uint8_t>(),
                                 ws->hole_traced.as<uint8_t>(), ws->bbox.as<int4>(), ws->flags.as<uint8_t>(),
                                 reinterpret_cast<int *>(ws->bbox.as<int4>() + (size_t)B * H), ws->offs.as<int>()));
  OCRB_TRY(launch_contour_count(ctx, n, ws->offs.as<int>()));
  int nc32 = 0;
  OCRB_TRY(read_scalar(ctx, ws->offs.as<int>() + contour_count_slot(n), &nc32));
  const int64_t nc = nc32;
  st->n_contours = nc;
  st->n_points = 0;
  if (nc == 0) return OCRB_OK;
  OCRB_TRY(ws->start_idx.reserve(nc * 8));
  OCRB_TRY(ws->kind.reserve(nc));
  OCRB_TRY(ws->lengths.reserve(nc * 4));
  OCRB_TRY(ws->chain_off.reserve((nc + 1) * 8));
  OCRB_TRY(ws->dp_count.reserve(nc * 4));
  OCRB_TRY(ws->scan_scratch.reserve(scan_scratch_elems(nc) * 8));
  OCRB_TRY(launch_contour_records(ctx, ws->flags.as<uint8_t>(), ws->offs.as<int>(), n, ws->start_idx.as<int64_t>(),
                                  ws->kind.as<uint8_t>()));
  OCRB_TRY(launch_trace_count(ctx, bitmap, H, W, ws->start_idx.as<int64_t>(), ws->kind.as<uint8_t>(), nc, ws->lengths.as<int>()));
  OCRB_TRY((exclusive_scan<int, int64_t>(ctx, ws->lengths.as<int>(), nc, ws->chain_off.as<int64_t>(),
                                         ws->scan_scratch.as<int64_t>())));
  int64_t np = 0;
  OCRB_TRY(read_scalar(ctx, ws->chain_off.as<int64_t>() + nc, &np));
  st->n_points = np;
  OCRB_TRY(ws->chain.reserve(np * 4));
  OCRB_TRY(launch_trace_store(ctx, bitmap, H, W, ws->start_idx.as<int64_t>(), ws->kind.as<uint8_t>(), nc,
                              ws->chain_off.as<int64_t>(), ws->chain.as<ushort2>()));
  if (with_dp) {
    OCRB_TRY(ws->dp_out.reserve(np * 4));
    OCRB_TRY(ws->stack.reserve(np * 4));
    OCRB_TRY(launch_approx_dp(ctx, ws->chain.as<ushort2>(), ws->chain_off.as<int64_t>(), nc, ws->stack.as<int>(),
                              ws->dp_out.as<ushort2>(), ws->dp_count.as<int>()));
  }
  return OCRB_OK;
}

// pred / bitmap on device.  adjust: device [B][2] f64.
static int run_postproc(ocrb_ctx *ctx, const float *pred, const uint8_t *bitmap, const double *adjust_dev, int B, int H,
                        int W, const ocrb_postproc_params &prm, ocrb_polygons *res) {
  PostprocWorkspace *ws = get_pp(ctx);
  const int64_t HW = (int64_t)H * W;
  res->n_images = B;
  res->image_offsets.assign(B + 1, 0);
  res->point_offsets.assign(1, 0);
  res->stats.assign((size_t)5 * B, 0);
  ContourStage cs;
  OCRB_TRY(run_contour_stage(ctx, ws, bitmap, B, H, W, &cs, true));
  const int64_t nc = cs.n_contours;
  OCRB_TRY(ws->stats.reserve((size_t)5 * B * 8));
  OCRB_TRY(ws->err.reserve(4));
  OCRB_CUDA(cudaMemsetAsync(ws->stats.p, 0, (size_t)5 * B * 8, ctx->stream));
  OCRB_CUDA(cudaMemsetAsync(ws->err.p, 0, 4, ctx->stream));
  int n_cand = 0;
  if (nc > 0) {
    OCRB_TRY(ws->cand_flag.reserve(nc));
    OCRB_TRY(ws->cand_rank.reserve((nc + 1) * 4));
    OCRB_TRY(launch_flag_ge4(ctx, ws->dp_count.as<int>(), nc, ws->cand_flag.as<uint8_t>()));
    OCRB_TRY((exclusive_scan<uint8_t, int>(ctx, ws->cand_flag.as<uint8_t>(), nc, ws->cand_rank.as<int>(), ws->scan_scratch.as<int>())));
    OCRB_TRY(read_scalar(ctx, ws->cand_rank.as<int>() + nc, &n_cand));
  }
  int n_kept = 0;
  int64_t n_kept_pts = 0;
  if (n_cand > 0) {
    OCRB_TRY(ws->cand_contour.reserve((size_t)n_cand * 4));
    OCRB_TRY(ws->scores.reserve((size_t)n_cand * 8));
    OCRB_TRY(ws->slab_units.reserve((size_t)n_cand * 8));
    OCRB_TRY(ws->slab_off.reserve((size_t)(n_cand + 1) * 8));
    OCRB_TRY(ws->out_count.reserve((size_t)n_cand * 4));
    OCRB_TRY(ws->status.reserve((size_t)n_cand));
    OCRB_TRY(ws->kept_flag.reserve((size_t)n_cand));
    OCRB_TRY(ws->kept_pts.reserve((size_t)n_cand * 4));
    OCRB_TRY(ws->kept_rank.reserve((size_t)(n_cand + 1) * 4));
    OCRB_TRY(ws->pt_off.reserve((size_t)(n_cand + 1) * 8));
    OCRB_TRY(ws->scan_scratch.reserve(scan_scratch_elems(n_cand) * 8));
    OCRB_TRY(launch_compact_index(ctx, ws->cand_flag.as<uint8_t>(), ws->cand_rank.as<int>(), nc, ws->cand_contour.as<int>()));
    // dims: the reference passes pred.get(0) = [H][W]: size[-2] = H, size[-1] = W
    OCRB_TRY(launch_box_score(ctx, pred, H, W, HW, ws->cand_contour.as<int>(), ws->start_idx.as<int64_t>(),
                              ws->chain_off.as<int64_t>(), ws->dp_out.as<ushort2>(), ws->dp_count.as<int>(), n_cand,
                              ws->scores.as<double>(), ws->err.as<int>()));
    OCRB_TRY(launch_unclip_slab_sizes(ctx, ws->cand_contour.as<int>(), ws->dp_count.as<int>(), n_cand, ws->slab_units.as<int64_t>()));
    OCRB_TRY((exclusive_scan<int64_t, int64_t>(ctx, ws->slab_units.as<int64_t>(), n_cand, ws->slab_off.as<int64_t>(),
                                               ws->scan_scratch.as<int64_t>())));
    int64_t total_units = 0;
    OCRB_TRY(read_scalar(ctx, ws->slab_off.as<int64_t>() + n_cand, &total_units));
    OCRB_TRY(ws->slabs.reserve((size_t)total_units * 8));
    OCRB_TRY(ws->cand_box.reserve((size_t)n_cand * 4 * sizeof(int2)));
    OCRB_TRY(launch_unclip(ctx, ws->cand_contour.as<int>(), ws->chain_off.as<int64_t>(), ws->dp_out.as<ushort2>(),
                           ws->dp_count.as<int>(), n_cand, ws->scores.as<double>(), prm.box_thresh, prm.min_size,
                           prm.unclip_factor, ws->slab_off.as<int64_t>(), ws->slabs.as<int2>(), ws->out_count.as<int>(),
                           ws->status.as<uint8_t>(), nullptr, ws->cand_box.as<int2>()));
    OCRB_TRY(launch_kept_sizes(ctx, ws->status.as<uint8_t>(), ws->out_count.as<int>(), n_cand, ws->kept_flag.as<uint8_t>(),
                               ws->kept_pts.as<int>()));
    OCRB_TRY((exclusive_scan<uint8_t, int>(ctx, ws->kept_flag.as<uint8_t>(), n_cand, ws->kept_rank.as<int>(), ws->scan_scratch.as<int>())));
    OCRB_TRY((exclusive_scan<int, int64_t>(ctx, ws->kept_pts.as<int>(), n_cand, ws->pt_off.as<int64_t>(), ws->scan_scratch.as<int64_t>())));
    OCRB_TRY(read_scalar(ctx, ws->kept_rank.as<int>() + n_cand, &n_kept));
    OCRB_TRY(read_scalar(ctx, ws->pt_off.as<int64_t>() + n_cand, &n_kept_pts));
  }
  OCRB_TRY(launch_stats(ctx, ws->start_idx.as<int64_t>(), ws->dp_count.as<int>(), nc, HW, ws->cand_contour.as<int>(), n_cand,
                        ws->status.as<uint8_t>(), ws->stats.as<unsigned long long>()));
  std::vector<int> img;
  if (n_kept > 0) {
    OCRB_TRY(ws->out_xy.reserve((size_t)n_kept_pts * 8));
    OCRB_TRY(ws->out_scores.reserve((size_t)n_kept * 8));
    OCRB_TRY(ws->out_pt_off.reserve((size_t)n_kept * 8));
    OCRB_TRY(ws->out_image.reserve((size_t)n_kept * 4));
    OCRB_TRY(ws->out_box.reserve((size_t)n_kept * 4 * sizeof(int2)));
    OCRB_TRY(launch_emit_polygons(ctx, ws->cand_contour.as<int>(), ws->dp_count.as<int>(), ws->start_idx.as<int64_t>(), HW,
                                  n_cand, ws->status.as<uint8_t>(), ws->kept_rank.as<int>(), ws->pt_off.as<int64_t>(),
                                  ws->slab_off.as<int64_t>(), ws->slabs.as<int2>(), ws->out_count.as<int>(),
                                  ws->scores.as<double>(), adjust_dev, ws->out_xy.as<uint32_t>(), ws->out_scores.as<double>(),
                                  ws->out_pt_off.as<int64_t>(), ws->out_image.as<int>(), ws->cand_box.as<int2>(), ws->out_box.as<int2>()));
    res->xy.resize((size_t)n_kept_pts * 2);
    res->scores.resize(n_kept);
    res->point_offsets.resize(n_kept + 1);
    img.resize(n_kept);
    OCRB_CUDA(cudaMemcpyAsync(res->xy.data(), ws->out_xy.p, (size_t)n_kept_pts * 8, cudaMemcpyDeviceToHost, ctx->stream));
    OCRB_CUDA(cudaMemcpyAsync(res->scores.data(), ws->out_scores.p, (size_t)n_kept * 8, cudaMemcpyDeviceToHost, ctx->stream));
    OCRB_CUDA(cudaMemcpyAsync(res->point_offsets.data(), ws->out_pt_off.p, (size_t)n_kept * 8, cudaMemcpyDeviceToHost, ctx->stream));
    OCRB_CUDA(cudaMemcpyAsync(img.data(), ws->out_image.p, (size_t)n_kept * 4, cudaMemcpyDeviceToHost, ctx->stream));
  }
  std::vector<unsigned long long> st((size_t)5 * B);
  int err = 0;
  OCRB_CUDA(cudaMemcpyAsync(st.data(), ws->stats.p, (size_t)5 * B * 8, cudaMemcpyDeviceToHost, ctx->stream));
  OCRB_CUDA(cudaMemcpyAsync(&err, ws->err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  OCRB_TRY(sync(ctx));
  if (err) {  // cannot happen for maps up to 65535 px wide (run_contour_stage checks that); kept as an internal check
    set_error("post-processing: a candidate's bounding box does not fit the mask band (flags %d)", err);
    return OCRB_ERR_INTERNAL;
  }
  for (size_t i = 0; i < st.size(); ++i) res->stats[i] = (int64_t)st[i];
  res->point_offsets[n_kept] = n_kept_pts;
  for (int k = 0; k < n_kept; ++k) res->image_offsets[img[k] + 1] += 1;
  for (int b = 0; b < B; ++b) res->image_offsets[b + 1] += res->image_offsets[b];
  return OCRB_OK;
}

// pipeline.cu: the kept polygons' min-area boxes and image indices of the LAST post-processing call on this ctx
// (device arrays, in result order; valid until the next call)
void postproc_kept_boxes(ocrb_ctx *ctx, const int2 **boxes, const int **image) {
  PostprocWorkspace *ws = get_pp(ctx);
  *boxes = ws->out_box.as<int2>();
  *image = ws->out_image.as<int>();
}

// pipeline.cu entry: pred / bitmap / adjust already on the device; appends nothing, fills `res`
int postproc_device(ocrb_ctx *ctx, const float *pred_dev, const uint8_t *bitmap_dev, const double *adjust_dev, int B, int H,
                    int W, const ocrb_postproc_params &prm, ocrb_polygons *res) {
  return run_postproc(ctx, pred_dev, bitmap_dev, adjust_dev, B, H, W, prm, res);
}

// appends the polygons of `src` (a later chunk of images) to `dst`
void polygons_append(ocrb_polygons *dst, const ocrb_polygons *src) {
  if (dst->n_images == 0 && dst->image_offsets.empty()) {
    dst->image_offsets.assign(1, 0);
    dst->point_offsets.assign(1, 0);
  }
  const int64_t poly_base = dst->image_offsets.back(), pt_base = dst->point_offsets.back();
  for (int b = 0; b < src->n_images; ++b) dst->image_offsets.push_back(poly_base + src->image_offsets[b + 1]);
  for (size_t p = 1; p < src->point_offsets.size(); ++p) dst->point_offsets.push_back(pt_base + src->point_offsets[p]);
  dst->xy.insert(dst->xy.end(), src->xy.begin(), src->xy.end());
  dst->scores.insert(dst->scores.end(), src->scores.begin(), src->scores.end());
  dst->stats.insert(dst->stats.end(), src->stats.begin(), src->stats.end());
  dst->glyph_classes.insert(dst->glyph_classes.end(), src->glyph_classes.begin(), src->glyph_classes.end());
  if (src->glyphs_per_polygon) dst->glyphs_per_polygon = src->glyphs_per_polygon;
  dst->n_images += src->n_images;
}
ocrb_polygons *polygons_new() { return new ocrb_polygons(); }

// pipeline.cu: classes of the glyph tiles, group by group (host copies), in result order
void polygons_set_glyph_classes(ocrb_polygons *p, int k, const std::vector<PinBuf> &cls, const std::vector<int64_t> &kept) {
  p->glyphs_per_polygon = k;
  p->glyph_classes.clear();
  for (size_t g = 0; g < kept.size(); ++g) {
    const int32_t *src = cls[g].as<int32_t>();
    p->glyph_classes.insert(p->glyph_classes.end(), src, src + kept[g] * k);
  }
}

}  // namespace ocrb

using namespace ocrb;

extern "C" {

void ocrb_postproc_default_params(ocrb_postproc_params *p) {
  if (!p) return;
  p->thresh = 0.6;         // metrics.rs:38
  p->box_thresh = 0.7;     // metrics.rs:64
  p->min_size = 5.0;       // metrics.rs:66
  p->unclip_factor = 2.0;  // metrics.rs:103
}

int ocrb_get_boxes_and_box_scores(ocrb_ctx *ctx, const float *pred, const double *adjust, int B, int H, int W,
                                  const ocrb_postproc_params *params, ocrb_polygons **out) {
  OCRB_REQUIRE(ctx && pred && adjust && out, "null argument");
  OCRB_REQUIRE(B > 0 && H > 0 && W > 0, "bad shape B=%d H=%d W=%d", B, H, W);
  OCRB_CUDA(cudaSetDevice(ctx->device));
  ocrb_postproc_params prm;
  ocrb_postproc_default_params(&prm);
  if (params) prm = *params;
  const int64_t n = (int64_t)B * H * W;
  PostprocWorkspace *ws = get_pp(ctx);
  const void *pred_dev = nullptr;
  OCRB_TRY(to_device(ctx, 0, pred, (size_t)n * 4, &pred_dev));
  OCRB_TRY(ws->adjust.reserve((size_t)B * 16));
  OCRB_CUDA(cudaMemcpyAsync(ws->adjust.p, adjust, (size_t)B * 16, cudaMemcpyDefault, ctx->stream));
  OCRB_TRY(ws->bitmap.reserve(n));
  OCRB_TRY(launch_binarize(ctx, (const float *)pred_dev, n, (float)prm.thresh, ws->bitmap.as<uint8_t>()));
  ocrb_polygons *res = new ocrb_polygons();
  int rc = run_postproc(ctx, (const float *)pred_dev, ws->bitmap.as<uint8_t>(), ws->adjust.as<double>(), B, H, W, prm, res);
  if (rc != OCRB_OK) {
    delete res;
    cudaStreamSynchronize(ctx->stream);
    return rc;
  }
  *out = res;
  return OCRB_OK;
}

int ocrb_get_polygons_from_bitmap(ocrb_ctx *ctx, const float *pred, const uint8_t *bitmap, const double *adjust, int H, int W,
                                  const ocrb_postproc_params *params, ocrb_polygons **out) {
  OCRB_REQUIRE(ctx && pred && bitmap && adjust && out, "null argument");
  OCRB_REQUIRE(H > 0 && W > 0, "bad shape H=%d W=%d", H, W);
  OCRB_CUDA(cudaSetDevice(ctx->device));
  ocrb_postproc_params prm;
  ocrb_postproc_default_params(&prm);
  if (params) prm = *params;
  const int64_t n = (int64_t)H * W;
  PostprocWorkspace *ws = get_pp(ctx);
  const void *pred_dev = nullptr, *bm_dev = nullptr;
  OCRB_TRY(to_device(ctx, 0, pred, (size_t)n * 4, &pred_dev));
  OCRB_TRY(to_device(ctx, 1, bitmap, (size_t)n, &bm_dev));
  OCRB_TRY(ws->adjust.reserve(16));
  OCRB_CUDA(cudaMemcpyAsync(ws->adjust.p, adjust, 16, cudaMemcpyDefault, ctx->stream));
  ocrb_polygons *res = new ocrb_polygons();
  int rc = run_postproc(ctx, (const float *)pred_dev, (const uint8_t *)bm_dev, ws->adjust.as<double>(), 1, H, W, prm, res);
  if (rc != OCRB_OK) {
    delete res;
    cudaStreamSynchronize(ctx->stream);
    return rc;
  }
  *out = res;
  return OCRB_OK;
}

int ocrb_polygons_num_images(const ocrb_polygons *p) { return p ? p->n_images : 0; }
const int64_t *ocrb_polygons_image_offsets(const ocrb_polygons *p) { return p->image_offsets.data(); }
const int64_t *ocrb_polygons_point_offsets(const ocrb_polygons *p) { return p->point_offsets.data(); }
const uint32_t *ocrb_polygons_xy(const ocrb_polygons *p) { return p->xy.data(); }
const double *ocrb_polygons_scores(const ocrb_polygons *p) { return p->scores.data(); }
const int64_t *ocrb_polygons_stats(const ocrb_polygons *p) { return p->stats.data(); }
int ocrb_polygons_glyphs_per_polygon(const ocrb_polygons *p) { return p ? p->glyphs_per_polygon : 0; }
const int32_t *ocrb_polygons_glyph_classes(const ocrb_polygons *p) { return p->glyph_classes.data(); }
void ocrb_polygons_free(ocrb_polygons *p) { delete p; }

// ---- fine-grained hooks ------------------------------------------------------------------
int ocrb_find_contours(ocrb_ctx *ctx, const uint8_t *bitmap, int H, int W, int64_t *offsets, uint8_t *types,
                       int64_t contour_cap, int32_t *xy, int64_t point_cap, int64_t *n_contours, int64_t *n_points) {
  OCRB_REQUIRE(ctx && bitmap && n_contours && n_points, "null argument");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  PostprocWorkspace *ws = get_pp(ctx);
  const void *bm_dev = nullptr;
  OCRB_TRY(to_device(ctx, 1, bitmap, (size_t)H * W, &bm_dev));
  ContourStage cs;
  OCRB_TRY(run_contour_stage(ctx, ws, (const uint8_t *)bm_dev, 1, H, W, &cs, false));
  *n_contours = cs.n_contours;
  *n_points = cs.n_points;
  if (!offsets && !types && !xy) return sync(ctx);
  if (cs.n_contours > contour_cap || cs.n_points > point_cap) {
    set_error("find_contours: need %lld contours / %lld points", (long long)cs.n_contours, (long long)cs.n_points);
    return OCRB_ERR_CAPACITY;
  }
  if (cs.n_contours == 0) {
    if (offsets) offsets[0] = 0;
    return sync(ctx);
  }
  std::vector<ushort2> pts((size_t)cs.n_points);
  if (offsets) OCRB_CUDA(cudaMemcpyAsync(offsets, ws->chain_off.p, (size_t)(cs.n_contours + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (types) OCRB_CUDA(cudaMemcpyAsync(types, ws->kind.p, (size_t)cs.n_contours, cudaMemcpyDeviceToHost, ctx->stream));
  OCRB_CUDA(cudaMemcpyAsync(pts.data(), ws->chain.p, (size_t)cs.n_points * 4, cudaMemcpyDeviceToHost, ctx->stream));
  OCRB_TRY(sync(ctx));
  if (xy)
    for (int64_t i = 0; i < cs.n_points; ++i) { xy[2 * i] = pts[i].x; xy[2 * i + 1] = pts[i].y; }
  return OCRB_OK;
}

static int upload_points_u16(ocrb_ctx *ctx, PostprocWorkspace *ws, const int32_t *xy, int64_t n) {
  std::vector<ushort2> p((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    if (xy[2 * i] < 0 || xy[2 * i] > 65535 || xy[2 * i + 1] < 0 || xy[2 * i + 1] > 65535) {
      set_error("point %lld out of the u16 range", (long long)i);
      return OCRB_ERR_INVALID;
    }
    p[i] = make_ushort2((unsigned short)xy[2 * i], (unsigned short)xy[2 * i + 1]);
  }
  OCRB_TRY(ws->chain.reserve((size_t)n * 4));
  OCRB_TRY(ws->dp_out.reserve((size_t)n * 4));
  OCRB_TRY(ws->stack.reserve((size_t)n * 4));
  OCRB_TRY(ws->chain_off.reserve(16));
  OCRB_TRY(ws->dp_count.reserve(4));
  int64_t off[2] = {0, n};
  OCRB_CUDA(cudaMemcpyAsync(ws->chain.p, p.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  OCRB_CUDA(cudaMemcpyAsync(ws->chain_off.p, off, 16, cudaMemcpyHostToDevice, ctx->stream));
  return sync(ctx);  // p and off are stack/heap temporaries
}

int ocrb_approx_polygon(ocrb_ctx *ctx, const int32_t *chain_xy, int64_t n_pts, int32_t *out_xy, int64_t out_cap_pts, int64_t *n_out) {
  OCRB_REQUIRE(ctx && chain_xy && n_out && n_pts > 0, "bad argument");
  OCRB_REQUIRE(!is_device_ptr(chain_xy), "ocrb_approx_polygon takes host points");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  PostprocWorkspace *ws = get_pp(ctx);
  OCRB_TRY(upload_points_u16(ctx, ws, chain_xy, n_pts));
  OCRB_TRY(launch_approx_dp(ctx, ws->chain.as<ushort2>(), ws->chain_off.as<int64_t>(), 1, ws->stack.as<int>(),
                            ws->dp_out.as<ushort2>(), ws->dp_count.as<int>()));
  int m = 0;
  OCRB_TRY(read_scalar(ctx, ws->dp_count.as<int>(), &m));
  *n_out = m;
  if (m > out_cap_pts) { set_error("approx_polygon: need %d points", m); return OCRB_ERR_CAPACITY; }
  std::vector<ushort2> p((size_t)std::max(m, 1));
  if (m > 0) {
    OCRB_CUDA(cudaMemcpyAsync(p.data(), ws->dp_out.p, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    OCRB_TRY(sync(ctx));
    for (int i = 0; i < m; ++i) { out_xy[2 * i] = p[i].x; out_xy[2 * i + 1] = p[i].y; }
  }
  return OCRB_OK;
}

// uploads a polygon as a one-contour "DP result" so the batch kernels can be reused
static int upload_polygon(ocrb_ctx *ctx, PostprocWorkspace *ws, const int32_t *xy, int n) {
  OCRB_TRY(upload_points_u16(ctx, ws, xy, n));
  OCRB_CUDA(cudaMemcpyAsync(ws->dp_out.p, ws->chain.p, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  OCRB_CUDA(cudaMemcpyAsync(ws->dp_count.p, &n, 4, cudaMemcpyHostToDevice, ctx->stream));
  return sync(ctx);
}

int ocrb_box_score_fast(ocrb_ctx *ctx, const float *pred, int dim_m2, int dim_m1, const int32_t *xy, int n_pts, double *score) {
  OCRB_REQUIRE(ctx && pred && xy && score && n_pts > 0 && dim_m1 > 0 && dim_m2 > 0, "bad argument");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  PostprocWorkspace *ws = get_pp(ctx);
  const void *pred_dev = nullptr;
  OCRB_TRY(to_device(ctx, 0, pred, (size_t)dim_m1 * dim_m2 * 4, &pred_dev));
  OCRB_TRY(upload_polygon(ctx, ws, xy, n_pts));
  OCRB_TRY(ws->scores.reserve(8));
  OCRB_TRY(ws->err.reserve(4));
  OCRB_CUDA(cudaMemsetAsync(ws->err.p, 0, 4, ctx->stream));
  OCRB_TRY(launch_box_score(ctx, (const float *)pred_dev, dim_m2, dim_m1, (int64_t)dim_m1 * dim_m2, nullptr, nullptr,
                            ws->chain_off.as<int64_t>(), ws->dp_out.as<ushort2>(), ws->dp_count.as<int>(), 1,
                            ws->scores.as<double>(), ws->err.as<int>()));
  int err = 0;
  OCRB_CUDA(cudaMemcpyAsync(&err, ws->err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  OCRB_TRY(read_scalar(ctx, ws->scores.as<double>(), score));
  if (err) { set_error("box_score_fast capacity flags %d", err); return OCRB_ERR_CAPACITY; }
  return OCRB_OK;
}

static int run_single_unclip(ocrb_ctx *ctx, PostprocWorkspace *ws, const int32_t *xy, int n, double factor, double min_size,
                             int *n_out, std::vector<int2> *pts, double *sside, int2 *box) {
  OCRB_TRY(upload_polygon(ctx, ws, xy, n));
  int64_t units = 0;
  {
    int64_t cap = unclip_cap_h(n);
    units = (n + 1) + (3 * (int64_t)n + 3) + cap + cap + 2 * cap + 2 * (cap + 1);
  }
  OCRB_TRY(ws->slabs.reserve((size_t)units * 8));
  OCRB_TRY(ws->slab_off.reserve(16));
  OCRB_TRY(ws->scores.reserve(8));
  OCRB_TRY(ws->out_count.reserve(4));
  OCRB_TRY(ws->status.reserve(4));
  OCRB_TRY(ws->kept_pts.reserve(64 + 8));  // sside (8) + box (32)
  int64_t off[2] = {0, units};
  double one = 1.0;
  OCRB_CUDA(cudaMemcpyAsync(ws->slab_off.p, off, 16, cudaMemcpyHostToDevice, ctx->stream));
  OCRB_CUDA(cudaMemcpyAsync(ws->scores.p, &one, 8, cudaMemcpyHostToDevice, ctx->stream));
  OCRB_TRY(sync(ctx));
  double *sside_dev = ws->kept_pts.as<double>();
  int2 *box_dev = reinterpret_cast<int2 *>(ws->kept_pts.as<char>() + 16);
  OCRB_CUDA(cudaMemsetAsync(ws->kept_pts.p, 0, 64, ctx->stream));
  OCRB_TRY(launch_unclip(ctx, nullptr, ws->chain_off.as<int64_t>(), ws->dp_out.as<ushort2>(), ws->dp_count.as<int>(), 1,
                         ws->scores.as<double>(), 0.0, min_size, factor, ws->slab_off.as<int64_t>(), ws->slabs.as<int2>(),
                         ws->out_count.as<int>(), ws->status.as<uint8_t>(), sside_dev, box_dev));
  int ne = 0;
  OCRB_TRY(read_scalar(ctx, ws->out_count.as<int>(), &ne));
  *n_out = ne;
  if (pts && ne > 0) {
    pts->resize(ne);
    int64_t cap = unclip_cap_h(n);
    const int2 *src = ws->slabs.as<int2>() + (n + 1) + (3 * (int64_t)n + 3) + cap;
    OCRB_CUDA(cudaMemcpyAsync(pts->data(), src, (size_t)ne * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (sside) OCRB_CUDA(cudaMemcpyAsync(sside, sside_dev, 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (box) OCRB_CUDA(cudaMemcpyAsync(box, box_dev, 32, cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}

int ocrb_debug_approx_polygon_host(const int32_t *chain_xy, int64_t n_pts, int32_t *out_xy, int64_t out_cap_pts, int64_t *n_out) {
  OCRB_REQUIRE(chain_xy && out_xy && n_out && n_pts > 0 && n_pts < (1ll << 30) && out_cap_pts > 0, "bad argument");
  return ocrb::approx_polygon_host(chain_xy, n_pts, out_xy, out_cap_pts, n_out);
}

int ocrb_debug_min_area_bounding_box_host(const int32_t *xy, int n_pts, int32_t *box_xy, double *sside) {
  OCRB_REQUIRE(xy && box_xy && sside && n_pts > 0, "bad argument");
  return ocrb::min_area_bounding_box_host(xy, n_pts, box_xy, sside);
}

int ocrb_clip_polygon(const int32_t *xy, int n_pts, double factor, int shrink, int32_t *out_xy, int out_cap_pts, int *n_out, double *distance) {
  OCRB_REQUIRE(xy && out_xy && n_out && n_pts > 0 && out_cap_pts > 0, "bad argument");
  return ocrb::clip_polygon_host(xy, n_pts, factor, shrink, out_xy, out_cap_pts, n_out, distance);
}

int ocrb_expand_polygon(ocrb_ctx *ctx, const int32_t *xy, int n_pts, double factor, int32_t *out_xy, int out_cap_pts, int *n_out) {
  OCRB_REQUIRE(ctx && xy && n_out && n_pts > 0, "bad argument");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  PostprocWorkspace *ws = get_pp(ctx);
  std::vector<int2> pts;
  int ne = 0;
  OCRB_TRY(run_single_unclip(ctx, ws, xy, n_pts, factor, 0.0, &ne, &pts, nullptr, nullptr));
  *n_out = ne;
  if (ne > out_cap_pts) { set_error("expand_polygon: need %d points", ne); return OCRB_ERR_CAPACITY; }
  for (int i = 0; i < ne; ++i) { out_xy[2 * i] = pts[i].x; out_xy[2 * i + 1] = pts[i].y; }
  return OCRB_OK;
}

int ocrb_min_area_bounding_box(ocrb_ctx *ctx, const int32_t *xy, int n_pts, int32_t *box_xy, double *sside) {
  OCRB_REQUIRE(ctx && xy && box_xy && sside && n_pts > 0, "bad argument");
  OCRB_CUDA(cudaSetDevice(ctx->device));
  PostprocWorkspace *ws = get_pp(ctx);
  // points may be negative here (expanded polygons): upload as int2 directly
  OCRB_TRY(ws->slabs.reserve((size_t)n_pts * 8 + (size_t)(2 * n_pts + 2) * 16 + 64));
  int2 *pts_dev = ws->slabs.as<int2>();
  double2 *work = reinterpret_cast<double2 *>(ws->slabs.as<char>() + (((size_t)n_pts * 8 + 15) / 16) * 16);
  double2 *hull = work + n_pts;
  OCRB_TRY(ws->kept_pts.reserve(64 + 8));
  double *sside_dev = ws->kept_pts.as<double>();
  int2 *box_dev = reinterpret_cast<int2 *>(ws->kept_pts.as<char>() + 16);
  OCRB_CUDA(cudaMemcpyAsync(pts_dev, xy, (size_t)n_pts * 8, cudaMemcpyHostToDevice, ctx->stream));
  OCRB_TRY(launch_minrect_hook(ctx, pts_dev, n_pts, work, hull, box_dev, sside_dev));
  OCRB_CUDA(cudaMemcpyAsync(sside, sside_dev, 8, cudaMemcpyDeviceToHost, ctx->stream));
  OCRB_CUDA(cudaMemcpyAsync(box_xy, box_dev, 32, cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}

}  // extern "C"
