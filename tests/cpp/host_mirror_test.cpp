// Host-side C++ mirror of ocr-rs's modules (include/ocrb.hpp) exercised from compiled code: compiled and run by
// tests/test_cpp_host.py.  Without a CUDA device it checks the host-only helpers and the error behaviour (no CPU
// fallback); with one it runs the known-answer tests of the reference (metrics.rs:406-508) through the mirror.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "ocrb.hpp"

#define REQUIRE(c)                                                          \
  do {                                                                      \
    if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } \
  } while (0)

int main(int argc, char **argv) {
  using namespace ocr_rs;
  REQUIRE(ocrb_version() == OCRB_VERSION);
  if (argc > 1) {
    // vs.load(file) (text_detection/mod.rs:40-44): the native reader of tch's VarStore archives, on the fixture written
    // by libtorch's own serializer (tests/golden/varstore_libtorch.ot: 18 tensors, conv1.weight is [4][1][7][7])
    ocrb_varstore *vs = nullptr;
    REQUIRE(ocrb_varstore_open(argv[1], &vs) == OCRB_OK && ocrb_varstore_count(vs) == 18);
    bool found = false;
    for (int i = 0; i < ocrb_varstore_count(vs); ++i) {
      if (std::string(ocrb_varstore_name(vs, i)) != "conv1.weight") continue;
      const float *data = nullptr;
      const int64_t *shape = nullptr;
      int64_t numel = 0;
      int ndim = 0;
      REQUIRE(ocrb_varstore_tensor(vs, i, &data, &numel, &shape, &ndim) == OCRB_OK);
      found = data && numel == 4 * 49 && ndim == 4 && shape[0] == 4 && shape[1] == 1 && shape[2] == 7 && shape[3] == 7;
    }
    ocrb_varstore_close(vs);
    REQUIRE(found);
    REQUIRE(ocrb_varstore_open("/nonexistent/model.ot", &vs) != OCRB_OK);  // vs.load errors on a missing file
  }
  // utils::VALUES (utils.rs:7)
  REQUIRE(utils::class_to_char(0) == 'A' && utils::class_to_char(26) == 'a' && utils::class_to_char(52) == '0' && utils::class_to_char(62) == '?');
  {  // utils::topk (utils.rs:28-43) and utils::parse_dimensions (utils.rs:72-79)
    std::vector<double> p(62, 0.0);
    p[3] = 0.2; p[30] = 0.5; p[61] = 0.3;
    const auto t = utils::topk(p, 3);
    REQUIRE(t.size() == 3 && t[0].first == 'e' && t[0].second == 0.5 && t[1].first == '9' && t[2].first == 'D');
    bool threw = false;
    try { utils::topk(std::vector<double>(61, 0.0), 1); } catch (const Error &) { threw = true; }
    REQUIRE(threw);
    REQUIRE(utils::parse_dimensions("800x600") == (std::pair<uint32_t, uint32_t>(800, 600)));
    for (const char *bad : {"800", "800x600x3", "ax600", "800X600", "x600"}) {
      threw = false;
      try { utils::parse_dimensions(bad); } catch (const Error &) { threw = true; }
      REQUIRE(threw);
    }
  }
  {  // polygon::clip_polygon / shrink_polygon (polygon.rs:13-49) are host code: square +5 / rectangle shrunk by 0.75 A / P
    const auto ex = polygon::clip_polygon({{10, 10}, {20, 10}, {20, 20}, {10, 20}}, 2.0, polygon::OffsetType::Expand);
    REQUIRE(ex && ex->size() == 4 && (*ex)[0].x == 25 && (*ex)[0].y == 25 && (*ex)[2].x == 5 && (*ex)[2].y == 5);
    const auto sh = polygon::shrink_polygon({{0, 0}, {100, 0}, {100, 40}, {0, 40}}, 0.75);
    REQUIRE(sh && sh->size() == 4 && (*sh)[0].x == 89 && (*sh)[0].y == 29 && (*sh)[2].x == 11 && (*sh)[2].y == 11);
    REQUIRE(!polygon::clip_polygon({{5, 5}, {5, 5}, {5, 5}, {5, 5}}, 2.0, polygon::OffsetType::Expand));
  }
  int rw = 0, rh = 0;
  REQUIRE(ocrb_resize_dims(300, 200, 800, 800, &rw, &rh) == OCRB_OK && rw == 800 && rh == 533);  // image_ops.rs fixtures
  {  // evaluation metrics are host code: the reference's KATs (metrics.rs:648-678, :814-856) through the mirror
    const std::vector<Polygon> gt = {{{0, 0}, {10, 0}, {10, 10}, {0, 10}}, {{20, 20}, {30, 20}, {30, 30}, {20, 30}}};
    const std::vector<Polygon> pred = {{{1, 1}, {10, 0}, {10, 10}, {0, 10}}};
    const eval::MetricsItem m = eval::evaluate_image(gt, {false, false}, pred);
    REQUIRE(m.gt_care == 2 && m.det_care == 1 && m.det_matched == 1);
    REQUIRE(std::fabs(m.precision - 1.0) < 1e-15 && std::fabs(m.recall - 0.5) < 1e-15 && std::fabs(m.hmean - 0.6666666666666666) < 1e-15);
    const eval::MetricsItem all_ignored = eval::evaluate_image(gt, {true, true}, pred);
    REQUIRE(all_ignored.gt_care == 0 && all_ignored.det_care == 0 && all_ignored.precision == 1.0 && all_ignored.recall == 1.0);
    const std::vector<eval::MetricsItem> items = {{1., 0.5, 0.6666666666666666, 2, 1, 1}, {1., 1., 1., 0, 0, 0}, {1., 1., 1., 2, 2, 2}, {0.3333333333333333, 0.2, 0.25, 5, 3, 1}};
    const auto prh = eval::combine_results(items);
    REQUIRE(prh[0] == 0.6666666666666666 && prh[1] == 0.4444444444444444 && prh[2] == 0.5333333333333333);
    int64_t first = 0, count = 0;
    REQUIRE(ocrb_shard_range(1024, 7, 8, &first, &count) == OCRB_OK && first == 896 && count == 128);
  }
  int n_dev = 0;
  if (ocrb_device_count(&n_dev) != OCRB_OK || n_dev == 0) {
    bool threw = false;
    try {
      Context ctx(0);
    } catch (const Error &e) {
      threw = e.code == OCRB_ERR_CUDA;  // anyhow::Error in the reference; here the C code + message
    }
    REQUIRE(threw);
    std::puts("host-only checks ok (no CUDA device: Context refuses, no CPU fallback)");
    return 0;
  }
  Context ctx(0);
  // metrics.rs:486-508 binarize: strict > in f32 (0.6f > 0.6 is false)
  const float pred[6] = {0.1f, 0.6f, 0.6000001f, 0.7f, 0.59f, 1.0f};  // 0.6000001f is the next float above (float)0.6
  const std::vector<uint8_t> bm = text_detection::metrics::binarize(ctx, pred, 6, 0.6);
  REQUIRE(bm[0] == 0 && bm[1] == 0 && bm[2] == 1 && bm[3] == 1 && bm[4] == 0 && bm[5] == 1);
  // metrics.rs:406-424 get_min_area_bounding_box: axis-aligned rectangle, short side 5
  const std::vector<Point> rect = {{10, 10}, {30, 10}, {30, 15}, {10, 15}};
  auto [corners, sside] = text_detection::metrics::get_min_area_bounding_box(ctx, rect);
  REQUIRE(corners.size() == 4 && std::fabs(sside - 5.0) < 1e-9);
  // metrics.rs:150-184 box_score_fast: constant map -> its value
  std::vector<float> map(40 * 40, 0.25f);
  const double s = text_detection::metrics::box_score_fast(ctx, map.data(), 40, 40, rect);
  REQUIRE(std::fabs(s - 0.25) < 1e-12);
  // polygon.rs:51-56 expand_polygon: grows the rectangle; a degenerate polygon gives None
  auto grown = polygon::expand_polygon(ctx, rect, 2.0);
  REQUIRE(grown.has_value() && grown->size() >= 4);
  int minx = 1 << 30, maxx = -(1 << 30);
  for (const Point &p : *grown) { minx = p.x < minx ? p.x : minx; maxx = p.x > maxx ? p.x : maxx; }
  REQUIRE(minx < 10 && maxx > 30);
  // metrics.rs:37-56 on an empty map: no polygons, one (empty) entry per image
  std::vector<float> zeros(2 * 64 * 64, 0.0f);
  const double adj[4] = {1.0, 1.0, 1.0, 1.0};
  const PolygonScores ps = text_detection::metrics::get_boxes_and_box_scores(ctx, zeros.data(), adj, 2, 64, 64);
  REQUIRE(ps.polygons.size() == 2 && ps.polygons[0].empty() && ps.scores[1].empty());
  // a bright block is found again, scaled by adjust
  for (int y = 20; y < 40; ++y)
    for (int x = 8; x < 56; ++x) zeros[y * 64 + x] = 0.9f;
  const PolygonScores one = text_detection::metrics::get_boxes_and_box_scores(ctx, zeros.data(), adj, 2, 64, 64);
  REQUIRE(one.polygons[0].size() == 1 && one.polygons[1].empty() && one.scores[0][0] > 0.7);
  std::printf("device checks ok (%lld kernel launches)\n", (long long)ctx.launch_count());
  return 0;
}
