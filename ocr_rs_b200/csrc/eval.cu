// Evaluation metrics of the detector (metrics.rs:191-394): validate_measure / evaluate_image / combine_results and
// the polygon intersection-over-union they rest on.  Host code, like the reference's (geo-clipper on the CPU): it runs
// once per validation pass over a few polygons per image and is not on the device path.
//
// geo-clipper's intersection / union areas (metrics.rs:375-389; Clipper on integer coordinates, factor 1) are
// restated as exact region areas: the plane is cut into vertical slabs at every vertex and every edge crossing; inside
// a slab no two edges cross, so the edges of both polygons have a fixed vertical order and the region covered by both
// (non-zero winding each) is a set of trapezoids.  area(union) = region(A) + region(B) - area(A and B), with
// region(P) = area(P and P): for a simple polygon the shoelace area, for a self-touching / self-crossing one the area of
// the region Clipper's union covers.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace ocrb {

struct EPt { double x, y; };
struct EEdge { double x0, y0, x1, y1; int poly, dir; };  // x0 < x1; dir = +1 when the original edge ran towards +x

static double ring_area(const std::vector<EPt> &p) {
  double a = 0;
  for (size_t i = 0, n = p.size(); i < n; ++i) a += p[i].x * p[(i + 1) % n].y - p[(i + 1) % n].x * p[i].y;
  return std::fabs(a) * 0.5;
}

// area of {winding_A != 0} and {winding_B != 0}
static double intersection_area(const std::vector<EPt> &A, const std::vector<EPt> &B) {
  std::vector<EEdge> edges;
  std::vector<double> xs;
  const std::vector<EPt> *polys[2] = {&A, &B};
  for (int k = 0; k < 2; ++k) {
    const auto &P = *polys[k];
    for (size_t i = 0, n = P.size(); i < n; ++i) {
      EPt a = P[i], b = P[(i + 1) % n];
      xs.push_back(a.x);
      if (a.x == b.x) continue;  // vertical edges bound no area in a slab
      if (a.x < b.x) edges.push_back({a.x, a.y, b.x, b.y, k, +1});
      else edges.push_back({b.x, b.y, a.x, a.y, k, -1});
    }
  }
  // crossings between any two edges (either polygon may touch or cross itself)
  for (size_t i = 0; i < edges.size(); ++i)
    for (size_t j = i + 1; j < edges.size(); ++j) {
      const EEdge &e = edges[i], &f = edges[j];
      const double d1x = e.x1 - e.x0, d1y = e.y1 - e.y0, d2x = f.x1 - f.x0, d2y = f.y1 - f.y0;
      const double den = d1x * d2y - d1y * d2x;
      if (den == 0) continue;
      const double t = ((f.x0 - e.x0) * d2y - (f.y0 - e.y0) * d2x) / den;
      const double u = ((f.x0 - e.x0) * d1y - (f.y0 - e.y0) * d1x) / den;
      if (t > 0 && t < 1 && u > 0 && u < 1) xs.push_back(e.x0 + t * d1x);
    }
  std::sort(xs.begin(), xs.end());
  xs.erase(std::unique(xs.begin(), xs.end()), xs.end());
  double area = 0;
  struct Span { double ya, yb, ym; int poly, dir; };
  std::vector<Span> sp;
  for (size_t s = 0; s + 1 < xs.size(); ++s) {
    const double xa = xs[s], xb = xs[s + 1], xm = 0.5 * (xa + xb);
    sp.clear();
    for (const EEdge &e : edges) {
      if (!(e.x0 <= xa && e.x1 >= xb)) continue;
      const double k = (e.y1 - e.y0) / (e.x1 - e.x0);
      sp.push_back({e.y0 + k * (xa - e.x0), e.y0 + k * (xb - e.x0), e.y0 + k * (xm - e.x0), e.poly, e.dir});
    }
    std::sort(sp.begin(), sp.end(), [](const Span &p, const Span &q) { return p.ym < q.ym; });
    int w[2] = {0, 0};
    for (size_t i = 0; i + 1 < sp.size(); ++i) {
      w[sp[i].poly] += sp[i].dir;
      if (w[0] != 0 && w[1] != 0) area += 0.5 * ((sp[i + 1].ya - sp[i].ya) + (sp[i + 1].yb - sp[i].yb)) * (xb - xa);
    }
  }
  return area;
}

static std::vector<EPt> ring_from(const uint32_t *xy, int64_t n) {
  std::vector<EPt> p((size_t)n);
  for (int64_t i = 0; i < n; ++i) p[(size_t)i] = {(double)xy[2 * i], (double)xy[2 * i + 1]};
  return p;
}

}  // namespace ocrb

using namespace ocrb;

extern "C" {

int ocrb_polygon_iou(const uint32_t *a_xy, int n_a, const uint32_t *b_xy, int n_b, double *intersection, double *iou) {
  OCRB_REQUIRE(a_xy && b_xy && n_a >= 0 && n_b >= 0, "bad argument");
  try {
    const auto A = ring_from(a_xy, n_a), B = ring_from(b_xy, n_b);
    const double inter = intersection_area(A, B);
    const double uni = intersection_area(A, A) + intersection_area(B, B) - inter;
    if (intersection) *intersection = inter;
    if (iou) *iou = inter / uni;  // 0 / 0 = NaN like the reference's division (metrics.rs:388)
  } catch (const std::exception &e) {
    set_error("polygon_iou: %s", e.what());
    return OCRB_ERR_INTERNAL;
  }
  return OCRB_OK;
}

int ocrb_evaluate_image(const int64_t *gt_offsets, const uint32_t *gt_xy, int n_gt, const uint8_t *ignore_flags,
                        const int64_t *det_offsets, const uint32_t *det_xy, int n_det, ocrb_metrics_item *out) {
  OCRB_REQUIRE(out && n_gt >= 0 && n_det >= 0 && (n_gt == 0 || (gt_offsets && gt_xy && ignore_flags)) && (n_det == 0 || (det_offsets && det_xy)),
               "bad argument");
  try {
    const double area_precision_constraint = 0.5, iou_constraint = 0.5;  // metrics.rs:256-257
    std::vector<std::vector<EPt>> gt, det;
    std::vector<int> gt_dont_care, det_dont_care;
    for (int n = 0; n < n_gt; ++n) {
      gt.push_back(ring_from(gt_xy + 2 * gt_offsets[n], gt_offsets[n + 1] - gt_offsets[n]));
      if (ignore_flags[n]) gt_dont_care.push_back(n);
    }
    for (int n = 0; n < n_det; ++n) {
      det.push_back(ring_from(det_xy + 2 * det_offsets[n], det_offsets[n + 1] - det_offsets[n]));
      for (int dc : gt_dont_care) {  // metrics.rs:303-316
        const double inter = intersection_area(gt[dc], det.back());
        const double pd_area = ring_area(det.back());
        const double precision = pd_area == 0. ? 0. : inter / pd_area;
        if (precision > area_precision_constraint) {
          det_dont_care.push_back(n);
          break;
        }
      }
    }
    int64_t det_matched = 0;
    if (!gt.empty() && !det.empty()) {
      std::vector<int> gt_rect(gt.size(), 0), det_rect(det.size(), 0);
      std::vector<double> gt_region(gt.size()), det_region(det.size());
      for (size_t g = 0; g < gt.size(); ++g) gt_region[g] = intersection_area(gt[g], gt[g]);
      for (size_t d = 0; d < det.size(); ++d) det_region[d] = intersection_area(det[d], det[d]);
      auto in_gt_dont_care = [&](int v) { return std::find(gt_dont_care.begin(), gt_dont_care.end(), v) != gt_dont_care.end(); };
      for (int g = 0; g < (int)gt.size(); ++g)
        for (int d = 0; d < (int)det.size(); ++d) {
          const double inter = intersection_area(det[d], gt[g]);
          const double iou = inter / (det_region[d] + gt_region[g] - inter);
          // metrics.rs:329-334, kept literally: the detection index is looked up in the GROUND-TRUTH don't-care list
          if (gt_rect[g] == 0 && det_rect[d] == 0 && !in_gt_dont_care(g) && !in_gt_dont_care(d) && iou > iou_constraint) {
            gt_rect[g] = 1;
            det_rect[d] = 1;
            det_matched += 1;
          }
        }
    }
    const int64_t num_gt_care = (int64_t)gt.size() - (int64_t)gt_dont_care.size();
    const int64_t num_det_care = (int64_t)det.size() - (int64_t)det_dont_care.size();
    double recall, precision;
    if (num_gt_care == 0) {
      recall = 1.;
      precision = num_det_care > 0 ? 0. : 1.;
    } else {
      recall = (double)det_matched / (double)num_gt_care;
      precision = num_det_care == 0 ? 0. : (double)det_matched / (double)num_det_care;
    }
    const double hmean = precision + recall == 0. ? 0. : 2. * precision * recall / (precision + recall);
    *out = {precision, recall, hmean, num_gt_care, num_det_care, det_matched};
  } catch (const std::exception &e) {
    set_error("evaluate_image: %s", e.what());
    return OCRB_ERR_INTERNAL;
  }
  return OCRB_OK;
}

int ocrb_combine_results(const ocrb_metrics_item *items, int n, double *precision, double *recall, double *hmean) {
  OCRB_REQUIRE((items || n == 0) && n >= 0 && precision && recall && hmean, "bad argument");
  int64_t gt = 0, det = 0, matched = 0;
  for (int i = 0; i < n; ++i) {
    gt += items[i].gt_care;
    det += items[i].det_care;
    matched += items[i].det_matched;
  }
  const double r = gt != 0 ? (double)matched / (double)gt : 0.;
  const double p = det != 0 ? (double)matched / (double)det : 0.;
  *precision = p;
  *recall = r;
  *hmean = r + p != 0. ? 2. * (r * p) / (r + p) : 0.;
  return OCRB_OK;
}

}  // extern "C"
