// ocrb.hpp — C++ host side above the C ABI (include/ocrb.h), mirroring the module layout of lazareviczoran/ocr-rs.
//
// ocr-rs is compiled code (Rust) and no Rust toolchain exists in the build image, so the host-side mirror of its
// public functions is C++ (header-only, C++17): same module names, argument meaning and error behaviour
// (`anyhow::Result` -> exception ocr_rs::Error carrying the C error code and message; `Option::None` -> std::nullopt).
// Each function cites the reference function it stands for; INTEGRATION.md shows the equivalent Rust `ocrb-sys` shim.
// Pointers handed to these functions may be host or device memory (the library inspects them).
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ocrb.h"

namespace ocr_rs {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
  if (rc != OCRB_OK) throw Error(rc, std::string("libocrb error ") + std::to_string(rc) + ": " + ocrb_last_error());
}

// replaces the process-global `DEVICE` (main.rs:26-28): one context per (device, host thread)
class Context {
 public:
  explicit Context(int device = 0) { check(ocrb_ctx_create(device, &p_)); }
  ~Context() { if (p_) ocrb_ctx_destroy(p_); }
  Context(const Context &) = delete;
  Context &operator=(const Context &) = delete;
  Context(Context &&o) noexcept : p_(std::exchange(o.p_, nullptr)) {}
  ocrb_ctx *raw() const { return p_; }
  void synchronize() { check(ocrb_ctx_synchronize(p_)); }
  int64_t launch_count() const { return ocrb_ctx_launch_count(p_); }

 private:
  ocrb_ctx *p_ = nullptr;
};

struct Point { int32_t x, y; };
using Polygon = std::vector<std::array<uint32_t, 2>>;  // geo::Polygon<u32> exterior, no closing point

// metrics.rs:32-35
struct PolygonScores {
  std::vector<std::vector<Polygon>> polygons;  // Vec<MultiPolygon<u32>>: per image
  std::vector<std::vector<double>> scores;     // Vec<Vec<f64>>
  // detect_and_read only: classes of the glyph tiles cut from every polygon, [image][polygon][tile]
  std::vector<std::vector<std::vector<int32_t>>> glyph_classes;
};

namespace detail {
inline PolygonScores take(ocrb_polygons *h) {
  PolygonScores r;
  const int n = ocrb_polygons_num_images(h);
  const int64_t *io = ocrb_polygons_image_offsets(h), *po = ocrb_polygons_point_offsets(h);
  const uint32_t *xy = ocrb_polygons_xy(h);
  const double *sc = ocrb_polygons_scores(h);
  r.polygons.resize(n);
  r.scores.resize(n);
  r.glyph_classes.resize(n);
  const int gk = ocrb_polygons_glyphs_per_polygon(h);
  const int32_t *gc = gk > 0 ? ocrb_polygons_glyph_classes(h) : nullptr;
  for (int b = 0; b < n; ++b)
    for (int64_t p = io[b]; p < io[b + 1]; ++p) {
      Polygon poly;
      for (int64_t k = po[p]; k < po[p + 1]; ++k) poly.push_back({xy[2 * k], xy[2 * k + 1]});
      r.polygons[b].push_back(std::move(poly));
      r.scores[b].push_back(sc[p]);
      if (gc) r.glyph_classes[b].emplace_back(gc + p * gk, gc + (p + 1) * gk);
    }
  ocrb_polygons_free(h);
  return r;
}
inline std::vector<int32_t> flat(const std::vector<Point> &pts) {
  std::vector<int32_t> v;
  v.reserve(pts.size() * 2);
  for (const Point &p : pts) { v.push_back(p.x); v.push_back(p.y); }
  return v;
}
}  // namespace detail

namespace image_ops {
struct Preprocessed {
  std::vector<uint8_t> image;  // GrayImage, [h][w]
  double adjust_x, adjust_y;
};
// image_ops::preprocess_image (image_ops.rs:188-220) after the file decode: RGBA8 [src_h][src_w][4]
inline Preprocessed preprocess_image(Context &ctx, const uint8_t *rgba, int src_w, int src_h, std::pair<uint32_t, uint32_t> dims) {
  Preprocessed r;
  r.image.resize((size_t)dims.first * dims.second);
  check(ocrb_preprocess_rgba(ctx.raw(), rgba, src_w, src_h, (int)dims.first, (int)dims.second, r.image.data(), &r.adjust_x, &r.adjust_y));
  return r;
}
// image_ops::preprocess_image(file_path, target_dim) (image_ops.rs:188-220) from the ENCODED file bytes (JPEG / PNG): decode
// (image::open(..)?.into_rgba()) + resize + luma + pad in one library call; throws like the reference returns Err for a file
// it cannot decode
inline Preprocessed preprocess_image(Context &ctx, const std::vector<uint8_t> &file, std::pair<uint32_t, uint32_t> dims) {
  Preprocessed r;
  r.image.resize((size_t)dims.first * dims.second);
  const uint8_t *files[1] = {file.data()};
  const size_t sizes[1] = {file.size()};
  double adjust[2] = {0, 0};
  check(ocrb_preprocess_files(ctx.raw(), files, sizes, 1, (int)dims.first, (int)dims.second, r.image.data(), adjust));
  r.adjust_x = adjust[0];
  r.adjust_y = adjust[1];
  return r;
}
// image::open(file)?.into_luma() (image_ops.rs:78): decoded grey pixels [h][w]
inline std::vector<uint8_t> open_into_luma(Context &ctx, const std::vector<uint8_t> &file, int *w, int *h) {
  check(ocrb_image_info(file.data(), file.size(), w, h, nullptr));
  std::vector<uint8_t> luma((size_t)*w * *h);
  const uint8_t *files[1] = {file.data()};
  const size_t sizes[1] = {file.size()};
  const int64_t offs[1] = {0};
  check(ocrb_decode_images(ctx.raw(), files, sizes, 1, OCRB_PIXELS_LUMA, offs, luma.data()));
  return luma;
}
// image_ops::convert_image_to_tensor + to_kind(Float) (image_ops.rs:350-364)
inline std::vector<float> convert_image_to_tensor(Context &ctx, const uint8_t *image, int64_t n) {
  std::vector<float> t((size_t)n);
  check(ocrb_convert_image_to_tensor(ctx.raw(), image, n, t.data()));
  return t;
}
// image_ops::convert_tensor_to_image (image_ops.rs:367-381); `scale` as in text_detection/mod.rs:57
inline std::vector<uint8_t> convert_tensor_to_image(Context &ctx, const float *tensor, int64_t n, float scale = 1.0f) {
  std::vector<uint8_t> img((size_t)n);
  check(ocrb_convert_tensor_to_image(ctx.raw(), tensor, n, scale, img.data()));
  return img;
}
// image_ops::load_image_as_tensor (image_ops.rs:73-85) after the decode
inline std::vector<float> load_image_as_tensor(Context &ctx, const uint8_t *luma, int64_t n) {
  std::vector<float> t((size_t)n);
  check(ocrb_load_image_as_tensor(ctx.raw(), luma, n, t.data()));
  return t;
}
}  // namespace image_ops

namespace text_detection {
namespace model {
// resnet18(&vs.root()) + vs.load(model_file_path) (model.rs:154, text_detection/mod.rs:35-44)
class Resnet18 {
 public:
  Resnet18(Context &ctx, const std::string &model_file_path, int mode = OCRB_MODE_BF16) { check(ocrb_det_create_from_file(ctx.raw(), model_file_path.c_str(), mode, &p_)); }
  // from named OIHW float32 tensors (vs.variables())
  Resnet18(Context &ctx, const std::vector<std::string> &names, const std::vector<const float *> &data, const std::vector<int64_t> &numel,
           int mode = OCRB_MODE_BF16) {
    std::vector<const char *> cn;
    for (const auto &s : names) cn.push_back(s.c_str());
    check(ocrb_det_create(ctx.raw(), (int)cn.size(), cn.data(), data.data(), numel.data(), mode, &p_));
  }
  ~Resnet18() { if (p_) ocrb_det_destroy(p_); }
  Resnet18(const Resnet18 &) = delete;
  Resnet18 &operator=(const Resnet18 &) = delete;
  // net.forward_t(&images.view((b, 1, h, w)), false) (text_detection/mod.rs:52, :196): u8 grey levels in, f32 map out
  std::vector<float> forward_t(const uint8_t *images, int b, int h, int w) {
    std::vector<float> prob((size_t)b * h * w);
    check(ocrb_det_forward(p_, images, OCRB_U8, b, h, w, prob.data()));
    return prob;
  }
  void forward_t(const void *images, int dtype, int b, int h, int w, float *prob) { check(ocrb_det_forward(p_, images, dtype, b, h, w, prob)); }
  ocrb_det *raw() const { return p_; }

 private:
  ocrb_det *p_ = nullptr;
};
}  // namespace model

namespace metrics {
// metrics::binarize (metrics.rs:129-131)
inline std::vector<uint8_t> binarize(Context &ctx, const float *pred, int64_t n, double thresh) {
  std::vector<uint8_t> out((size_t)n);
  check(ocrb_binarize(ctx.raw(), pred, n, thresh, out.data()));
  return out;
}
// metrics::box_score_fast (metrics.rs:150-184): pred [dim_m2][dim_m1]
inline double box_score_fast(Context &ctx, const float *pred, int dim_m2, int dim_m1, const std::vector<Point> &points) {
  const std::vector<int32_t> xy = detail::flat(points);
  double s = 0.0;
  check(ocrb_box_score_fast(ctx.raw(), pred, dim_m2, dim_m1, xy.data(), (int)points.size(), &s));
  return s;
}
// metrics::get_min_area_bounding_box (metrics.rs:133-148) -> (4 corners, short side)
inline std::pair<std::vector<Point>, double> get_min_area_bounding_box(Context &ctx, const std::vector<Point> &points) {
  const std::vector<int32_t> xy = detail::flat(points);
  int32_t box[8];
  double sside = 0.0;
  check(ocrb_min_area_bounding_box(ctx.raw(), xy.data(), (int)points.size(), box, &sside));
  std::vector<Point> corners;
  for (int i = 0; i < 4; ++i) corners.push_back({box[2 * i], box[2 * i + 1]});
  return {corners, sside};
}
// metrics::get_polygons_from_bitmap (metrics.rs:58-127): one image; adjust = {adjust_x, adjust_y}
inline std::pair<std::vector<Polygon>, std::vector<double>> get_polygons_from_bitmap(Context &ctx, const float *pred, const uint8_t *bitmap,
                                                                                        const double adjust[2], int h, int w) {
  ocrb_polygons *out = nullptr;
  check(ocrb_get_polygons_from_bitmap(ctx.raw(), pred, bitmap, adjust, h, w, nullptr, &out));
  PolygonScores r = detail::take(out);
  return {std::move(r.polygons.at(0)), std::move(r.scores.at(0))};
}
// metrics::get_boxes_and_box_scores (metrics.rs:37-56): pred [b][h][w] (the reference's [b,1,h,w]), adjust [b][2]
inline PolygonScores get_boxes_and_box_scores(Context &ctx, const float *pred, const double *adjust_values, int b, int h, int w) {
  ocrb_polygons *out = nullptr;
  check(ocrb_get_boxes_and_box_scores(ctx.raw(), pred, adjust_values, b, h, w, nullptr, &out));
  return detail::take(out);
}
}  // namespace metrics
}  // namespace text_detection

namespace polygon {
// polygon::expand_polygon (polygon.rs:51-56): None when the offset is empty
inline std::optional<std::vector<Point>> expand_polygon(Context &ctx, const std::vector<Point> &points, double factor) {
  const std::vector<int32_t> xy = detail::flat(points);
  std::vector<int32_t> out(2 * (4 * points.size() + 16));
  int n = 0;
  check(ocrb_expand_polygon(ctx.raw(), xy.data(), (int)points.size(), factor, out.data(), (int)(out.size() / 2), &n));
  if (n == 0) return std::nullopt;
  std::vector<Point> r;
  for (int i = 0; i < n; ++i) r.push_back({out[2 * i], out[2 * i + 1]});
  return r;
}
// polygon::clip_polygon (polygon.rs:13-42) and shrink_polygon (polygon.rs:44-49): host code (no Context), as in the reference
enum class OffsetType { Shrink, Expand };
inline std::optional<std::vector<Point>> clip_polygon(const std::vector<Point> &points, double factor, OffsetType offset_type) {
  const std::vector<int32_t> xy = detail::flat(points);
  std::vector<int32_t> out(2 * (6 * points.size() + 32));
  int n = 0;
  check(ocrb_clip_polygon(xy.data(), (int)points.size(), factor, offset_type == OffsetType::Shrink ? 1 : 0, out.data(), (int)(out.size() / 2), &n, nullptr));
  if (n == 0) return std::nullopt;
  std::vector<Point> r;
  for (int i = 0; i < n; ++i) r.push_back({out[2 * i], out[2 * i + 1]});
  return r;
}
inline std::optional<std::vector<Point>> shrink_polygon(const std::vector<Point> &points, double factor) {
  return clip_polygon(points, factor, OffsetType::Shrink);
}
}  // namespace polygon

namespace utils {
// utils::VALUES (utils.rs:7)
inline const char *const VALUES = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789";
constexpr int VALUES_COUNT = 62;
inline char class_to_char(int cls) { return ocrb_class_to_char(cls); }
// utils::topk (utils.rs:28-43): the k largest of 62 class scores (run_prediction hands it softmax(-1, Double) of the
// logits) as (character, value), largest first; equal values keep the lower class first.  A vector of another length is
// the reference's panic on an unexpected tensor shape.
inline std::vector<std::pair<char, double>> topk(const std::vector<double> &scores, int k) {
  if ((int)scores.size() != VALUES_COUNT) throw Error(OCRB_ERR_INVALID, "unexpected tensor shape [" + std::to_string(scores.size()) + "]");
  if (k < 0 || k > VALUES_COUNT) throw Error(OCRB_ERR_INVALID, "k out of range");
  std::vector<int> order(VALUES_COUNT);
  for (int i = 0; i < VALUES_COUNT; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return scores[a] > scores[b]; });
  std::vector<std::pair<char, double>> r;
  for (int i = 0; i < k; ++i) r.emplace_back(VALUES[order[i]], scores[order[i]]);
  return r;
}
// utils::parse_dimensions ("800x800", utils.rs:72-79): exactly two 'x'-separated unsigned integers
inline std::pair<uint32_t, uint32_t> parse_dimensions(const std::string &dims) {
  const auto bad = [&]() { return Error(OCRB_ERR_INVALID, "Could not parse dimensions value: " + dims); };
  const size_t x = dims.find('x');
  if (x == std::string::npos || x == 0) throw bad();
  std::string a = dims.substr(0, x), b = dims.substr(x + 1);
  if (!b.empty() && b.back() == 'x') b.pop_back();  // split_terminator drops one trailing separator
  if (b.empty() || b.find('x') != std::string::npos) throw bad();
  for (const std::string *s : {&a, &b})
    for (char c : *s)
      if (c < '0' || c > '9') throw bad();
  if (a.size() > 10 || b.size() > 10) throw bad();
  const unsigned long long w = std::stoull(a), h = std::stoull(b);
  if (w > 0xFFFFFFFFull || h > 0xFFFFFFFFull) throw bad();
  return {(uint32_t)w, (uint32_t)h};
}
}  // namespace utils

namespace char_recognition {
namespace model {
// Net::new(&vs.root()) + vs.load(model_file_path) (char_recognition/model.rs:12-25, mod.rs:43-45)
class Net {
 public:
  Net(Context &ctx, const std::string &model_file_path) { check(ocrb_rec_create_from_file(ctx.raw(), model_file_path.c_str(), &p_)); }
  ~Net() { if (p_) ocrb_rec_destroy(p_); }
  Net(const Net &) = delete;
  Net &operator=(const Net &) = delete;
  // forward_t(xs, false): glyphs [b][784] f32 in [0, 1] -> logits [b][62]
  std::vector<float> forward_t(const float *glyphs, int b) {
    std::vector<float> logits((size_t)b * 62);
    check(ocrb_rec_forward(p_, glyphs, b, logits.data(), nullptr, nullptr));
    return logits;
  }
  // run_prediction's tail (mod.rs:53-56): softmax(-1, Double) + topk(1) -> (character, probability) per glyph
  std::vector<std::pair<char, double>> predict(const uint8_t *glyphs_u8, int b) {
    std::vector<int32_t> am((size_t)b);
    std::vector<double> pr((size_t)b);
    check(ocrb_rec_forward_u8(p_, glyphs_u8, b, nullptr, am.data(), pr.data()));
    std::vector<std::pair<char, double>> r;
    for (int i = 0; i < b; ++i) r.emplace_back(ocrb_class_to_char(am[i]), pr[i]);
    return r;
  }
  ocrb_rec *raw() const { return p_; }

 private:
  ocrb_rec *p_ = nullptr;
};
}  // namespace model
}  // namespace char_recognition

// run_text_detection's device part for a batch (text_detection/mod.rs:46-67, :188-204) + glyph classes in one call
inline PolygonScores detect_and_recognize(text_detection::model::Resnet18 &det, char_recognition::model::Net *rec, const uint8_t *images,
                                          const double *adjust_values, int b, int h, int w, const uint8_t *glyphs = nullptr, int n_glyphs = 0,
                                          int32_t *glyph_classes = nullptr) {
  ocrb_polygons *out = nullptr;
  check(ocrb_detect_and_recognize(det.raw(), rec ? rec->raw() : nullptr, images, adjust_values, b, h, w, nullptr, glyphs, n_glyphs, glyph_classes, &out));
  return detail::take(out);
}

// the same with the recognition net fed by the crop glue (README.md:20-26 "character segmentation"): glyphs_per_polygon
// tiles cut from every kept polygon's min-area rectangle; PolygonScores::glyph_classes holds their classes
inline PolygonScores detect_and_read(text_detection::model::Resnet18 &det, char_recognition::model::Net &rec, const uint8_t *images,
                                     const double *adjust_values, int b, int h, int w, int glyphs_per_polygon = 4) {
  ocrb_polygons *out = nullptr;
  check(ocrb_detect_and_read(det.raw(), rec.raw(), images, adjust_values, b, h, w, nullptr, glyphs_per_polygon, &out));
  return detail::take(out);
}

// evaluation metrics (metrics.rs:191-394); host code, no device needed
namespace eval {
using MetricsItem = ocrb_metrics_item;  // metrics.rs:22-30
namespace detail {
inline void csr(const std::vector<Polygon> &polys, std::vector<int64_t> &off, std::vector<uint32_t> &xy) {
  off.assign(1, 0);
  for (const Polygon &p : polys) {
    for (const auto &q : p) { xy.push_back(q[0]); xy.push_back(q[1]); }
    off.push_back((int64_t)xy.size() / 2);
  }
}
}  // namespace detail
// metrics.rs:251-372
inline MetricsItem evaluate_image(const std::vector<Polygon> &gt_points, const std::vector<bool> &ignore_flags, const std::vector<Polygon> &pred) {
  std::vector<int64_t> go, po;
  std::vector<uint32_t> gxy, pxy;
  detail::csr(gt_points, go, gxy);
  detail::csr(pred, po, pxy);
  std::vector<uint8_t> ig(ignore_flags.begin(), ignore_flags.end());
  MetricsItem m{};
  check(ocrb_evaluate_image(go.data(), gxy.data(), (int)gt_points.size(), ig.data(), po.data(), pxy.data(), (int)pred.size(), &m));
  return m;
}
// metrics.rs:191-218
inline std::vector<MetricsItem> validate_measure(const std::vector<std::vector<Polygon>> &polygons, const std::vector<std::vector<bool>> &ignore_tags,
                                                 const std::vector<std::vector<Polygon>> &pred, const std::vector<std::vector<double>> &scores) {
  std::vector<MetricsItem> out;
  for (size_t i = 0; i < polygons.size(); ++i) {
    std::vector<Polygon> kept;
    for (size_t k = 0; k < pred[i].size(); ++k)
      if (scores[i][k] >= 0.6) kept.push_back(pred[i][k]);
    out.push_back(evaluate_image(polygons[i], ignore_tags[i], kept));
  }
  return out;
}
// metrics.rs:229-249 -> (precision, recall, hmean)
inline std::array<double, 3> combine_results(const std::vector<MetricsItem> &results) {
  std::array<double, 3> r{};
  check(ocrb_combine_results(results.data(), (int)results.size(), &r[0], &r[1], &r[2]));
  return r;
}
}  // namespace eval

// several GPUs of one box from ONE process (text_detection/mod.rs:188-204 over ocrb_shards): model files are read natively
class Shards {
 public:
  Shards(const std::vector<int> &devices, const std::string &det_model_file, const std::string &rec_model_file, bool bf16 = true) {
    check(ocrb_shards_create_from_files(devices.data(), (int)devices.size(), det_model_file.c_str(), rec_model_file.empty() ? nullptr : rec_model_file.c_str(),
                                        bf16 ? OCRB_MODE_BF16 : OCRB_MODE_FP32, &p_));
  }
  ~Shards() { if (p_) ocrb_shards_destroy(p_); }
  Shards(const Shards &) = delete;
  Shards &operator=(const Shards &) = delete;
  int count() const { return ocrb_shards_count(p_); }
  PolygonScores detect_and_read(const uint8_t *images, const double *adjust_values, int b, int h, int w, int glyphs_per_polygon = 4) {
    ocrb_polygons *out = nullptr;
    check(ocrb_detect_and_read_sharded(p_, images, adjust_values, b, h, w, nullptr, glyphs_per_polygon, &out));
    return detail::take(out);
  }

 private:
  ocrb_shards *p_ = nullptr;
};

}  // namespace ocr_rs
