/*
 * ORACLE — test infrastructure only.  Never linked, imported or executed by the
 * product path (ocr_rs_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it.
 *
 * Plain-C, single-threaded restatement of the reference's detection post-processing
 *   /root/reference/src/text_detection/metrics.rs:37-184   (binarize, contours -> DP ->
 *       box score -> unclip -> min-area-rect filter -> rescale)
 *   /root/reference/src/polygon.rs:13-56                    (expand_polygon)
 *   /root/reference/src/image_ops.rs:188-220                (preprocess_image resize/luma/pad)
 * The arithmetic of those functions lives in third-party crates whose sources are not in
 * /root/reference (Cargo.lock pins): imageproc 0.22.0 (find_contours, arc_length,
 * approximate_polygon_dp, draw_polygon_mut, min_area_rect), image 0.23.11 (resize
 * Triangle, to_luma), geo 0.15.0 (unsigned_area, euclidean_length), geo-clipper
 * 0.4.1-alpha.0@54577fb -> clipper-sys 0.3.3-alpha.0@7f6d3a0 (Angus Johnson's Clipper 6:
 * ClipperOffset + union/pftPositive).  Their published algorithms are restated here
 *  and PINNED by the reference's own golden tests
 * (metrics.rs:406-646), reproduced by tests/test_oracle_goldens.py.
 *
 * Parity status: post-processing pinned (goldens).  Clipper's clean-up union is restated as
 * "boundary of the positive-winding region" (orc_union_positive) and pinned three ways: the
 * 4+4 golden polygons of metrics.rs:510-646; the reference's gt_shrinked_img{55,224,494,545}.png
 * and mask_img*.png regenerated bit-exactly from its ground-truth polygon files through
 * clip_polygon(Shrink) + draw_polygon_mut (image_ops.rs:222-277, tests/test_oracle_goldens.py);
 * and an independent winding-number rasteriser that shares no code with the walk
 * (oracle/region_check.c, tests/test_union_region.py).
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int32_t x, y; } ipt;

/* ------------------------------------------------------------------------------------
 * binarize — metrics.rs:129-131.  pred.gt(thresh) with the f64 scalar demoted to f32.
 * ---------------------------------------------------------------------------------- */
void orc_binarize(const float *pred, int64_t n, double thresh, uint8_t *out) {
  float t = (float)thresh;
  for (int64_t i = 0; i < n; ++i) out[i] = pred[i] > t ? 1 : 0;
}

/* ------------------------------------------------------------------------------------
 * find_contours — imageproc 0.22.0 contours.rs (Suzuki–Abe), called at metrics.rs:78-81.
 * Sequential, with the sign-marking work array, exactly as the crate does it
 * (SURVEY.md A.1).  Output: points of all contours concatenated, offsets[n+1],
 * types[n] (0 = Outer, 1 = Hole).  Returns number of contours, or -1 on capacity.
 * ---------------------------------------------------------------------------------- */
static const int RING_DX[8] = {-1, -1, 0, 1, 1, 1, 0, -1}; /* W NW N NE E SE S SW */
static const int RING_DY[8] = {0, -1, -1, -1, 0, 1, 1, 1};

static inline int dir_of(int dx, int dy) {
  for (int k = 0; k < 8; ++k)
    if (RING_DX[k] == dx && RING_DY[k] == dy) return k;
  return -1;
}

int orc_find_contours(const uint8_t *img, int W, int H, ipt *pts, int64_t max_pts,
                      int64_t *offsets, uint8_t *types, int max_contours) {
  int32_t *v = (int32_t *)malloc(sizeof(int32_t) * (size_t)W * H);
  for (int64_t i = 0; i < (int64_t)W * H; ++i) v[i] = img[i] > 0 ? 1 : 0;
#define V(x, y) v[(int64_t)(y) * W + (x)]
#define NZ(x, y) ((x) >= 0 && (y) >= 0 && (x) < W && (y) < H && V(x, y) != 0)
  int n = 0;
  int64_t np = 0;
  int32_t nbd = 1;
  offsets[0] = 0;
  for (int y = 0; y < H; ++y) {
    for (int x = 0; x < W; ++x) {
      if (V(x, y) == 0) continue;
      int from = -1, type = 0;
      if (V(x, y) == 1 && x > 0 && V(x - 1, y) == 0) { from = 0; type = 0; }          /* W */
      else if (V(x, y) > 0 && x + 1 < W && V(x + 1, y) == 0) { from = 4; type = 1; }  /* E */
      if (from < 0) continue;
      if (n >= max_contours) { free(v); return -1; }
      nbd += 1;
      /* clockwise search starting AT `from` */
      int d1 = -1;
      for (int k = 0; k < 8; ++k) {
        int d = (from + k) & 7;
        if (NZ(x + RING_DX[d], y + RING_DY[d])) { d1 = d; break; }
      }
      if (d1 < 0) {
        if (np + 1 > max_pts) { free(v); return -1; }
        pts[np].x = x; pts[np].y = y; np++;
        V(x, y) = -nbd;
      } else {
        int p1x = x + RING_DX[d1], p1y = y + RING_DY[d1];
        int p2x = p1x, p2y = p1y, p3x = x, p3y = y;
        for (;;) {
          if (np + 1 > max_pts) { free(v); return -1; }
          pts[np].x = p3x; pts[np].y = p3y; np++;
          int dp2 = dir_of(p2x - p3x, p2y - p3y);
          /* counter-clockwise search: start just CCW of dir(p2), end at dir(p2) itself */
          int d4 = -1, east_zero = 0;
          for (int k = 1; k <= 8; ++k) {
            int d = (dp2 - k) & 7;
            if (NZ(p3x + RING_DX[d], p3y + RING_DY[d])) { d4 = d; break; }
            if (d == 4) east_zero = 1;
          }
          /* d4 always exists: p2 itself is non-zero */
          int p4x = p3x + RING_DX[d4], p4y = p3y + RING_DY[d4];
          if (p3x + 1 == W || east_zero) V(p3x, p3y) = -nbd;
          else if (V(p3x, p3y) == 1) V(p3x, p3y) = nbd;
          if (p4x == x && p4y == y && p3x == p1x && p3y == p1y) break;
          p2x = p3x; p2y = p3y; p3x = p4x; p3y = p4y;
        }
      }
      types[n] = (uint8_t)type;
      n++;
      offsets[n] = np;
    }
  }
#undef V
#undef NZ
  free(v);
  return n;
}

/* ------------------------------------------------------------------------------------
 * arc_length(closed) — imageproc geometry.rs, called at metrics.rs:87
 * ---------------------------------------------------------------------------------- */
static double pt_dist(ipt a, ipt b) {
  double dx = (double)a.x - (double)b.x, dy = (double)a.y - (double)b.y;
  return sqrt(dx * dx + dy * dy);
}

double orc_arc_length(const ipt *c, int64_t n, int closed) {
  double len = 0.0;
  for (int64_t i = 0; i + 1 < n; ++i) len += pt_dist(c[i], c[i + 1]);
  if (n > 2 && closed) len += pt_dist(c[0], c[n - 1]);
  return len;
}

/* ------------------------------------------------------------------------------------
 * approximate_polygon_dp(curve, eps, closed=true) — imageproc geometry.rs, called at
 * metrics.rs:91.  Recursive, line through the two chain ends as
 * a = y0-y1, b = x1-x0, c = x0*y1 - x1*y0; distance |a x + b y + c| / sqrt(a^2+b^2);
 * FIRST index with strictly largest distance; closed => final pop().
 * ---------------------------------------------------------------------------------- */
static int64_t dp_rec(const ipt *c, int64_t lo, int64_t hi, double eps, ipt *out) {
  /* returns number of points written for the open chain c[lo..=hi] */
  double x0 = c[lo].x, y0 = c[lo].y, x1 = c[hi].x, y1 = c[hi].y;
  double a = y0 - y1, b = x1 - x0, cc = x0 * y1 - x1 * y0;
  double den = sqrt(a * a + b * b);
  double dmax = 0.0;
  int64_t index = lo;
  for (int64_t i = lo + 1; i <= hi; ++i) {
    double d = fabs(a * (double)c[i].x + b * (double)c[i].y + cc) / den;
    if (d > dmax) { index = i; dmax = d; }
  }
  if (dmax > eps) {
    int64_t n1 = dp_rec(c, lo, index, eps, out);
    n1 -= 1; /* partial1.pop() */
    int64_t n2 = dp_rec(c, index, hi, eps, out + n1);
    return n1 + n2;
  }
  out[0] = c[lo];
  out[1] = c[hi];
  return 2;
}

int64_t orc_approx_dp(const ipt *c, int64_t n, double eps, int closed, ipt *out) {
  if (n <= 0) return 0;
  int64_t m = dp_rec(c, 0, n - 1, eps, out);
  if (closed) m -= 1;
  return m;
}

/* ------------------------------------------------------------------------------------
 * box_score_fast — metrics.rs:150-184 with imageproc draw_polygon_mut.
 * `dim_m2` = size[-2] (reference calls it w, clamps x with it), `dim_m1` = size[-1]
 * (reference calls it h, clamps y with it) — D10 kept literally.
 * pred is [dim_m2, dim_m1] row-major f32.  Products in f32 (pred * {0,1}), sum in f64.
 * ---------------------------------------------------------------------------------- */
static inline int32_t clampi(int64_t v, int64_t lo, int64_t hi) {
  return (int32_t)(v < lo ? lo : (v > hi ? hi : v));
}

static int cmp_i32(const void *a, const void *b) {
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return (x > y) - (x < y);
}

/* draws mask (mw x mh, row-major bytes 0/1) */
static void draw_polygon_mask(uint8_t *mask, int mw, int mh, const ipt *poly, int n, uint8_t value) {
  if (n == 0) return;
  int32_t y_min = INT32_MAX, y_max = INT32_MIN;
  for (int i = 0; i < n; ++i) {
    if (poly[i].y < y_min) y_min = poly[i].y;
    if (poly[i].y > y_max) y_max = poly[i].y;
  }
  y_min = y_min < mh - 1 ? y_min : mh - 1; if (y_min < 0) y_min = 0;
  y_max = y_max < mh - 1 ? y_max : mh - 1; if (y_max < 0) y_max = 0;
  int32_t *xs = (int32_t *)malloc(sizeof(int32_t) * (size_t)(2 * n + 2));
  for (int32_t y = y_min; y <= y_max; ++y) {
    int k = 0;
    for (int e = 0; e < n; ++e) {
      ipt p0 = poly[e], p1 = poly[(e + 1) % n];
      if ((p0.y <= y && p1.y >= y) || (p1.y <= y && p0.y >= y)) {
        if (p0.y == p1.y) { xs[k++] = p0.x; xs[k++] = p1.x; }
        else if (p0.y == y || p1.y == y) {
          if (p1.y > y) xs[k++] = p0.x;
          if (p0.y > y) xs[k++] = p1.x;
        } else {
          float fraction = (float)(y - p0.y) / (float)(p1.y - p0.y);
          float inter = (float)p0.x + fraction * (float)(p1.x - p0.x);
          xs[k++] = (int32_t)roundf(inter);
        }
      }
    }
    qsort(xs, (size_t)k, sizeof(int32_t), cmp_i32);
    for (int c = 0; c + 1 < k; c += 2) {
      int32_t from = xs[c] < mw ? xs[c] : mw;
      int32_t to = xs[c + 1] < mw - 1 ? xs[c + 1] : mw - 1;
      if (from < mw && to >= 0) {
        if (from < 0) from = 0;
        if (to < 0) to = 0;
        for (int32_t x = from; x <= to; ++x) mask[(int64_t)y * mw + x] = value;
      }
    }
  }
  free(xs);
  /* outline: BresenhamLineIter on f32 end points (integers here) */
  for (int e = 0; e < n; ++e) {
    float x0 = (float)poly[e].x, y0 = (float)poly[e].y;
    float x1 = (float)poly[(e + 1) % n].x, y1 = (float)poly[(e + 1) % n].y;
    int steep = fabsf(y1 - y0) > fabsf(x1 - x0);
    if (steep) { float t = x0; x0 = y0; y0 = t; t = x1; x1 = y1; y1 = t; }
    if (x0 > x1) { float t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; }
    float dx = x1 - x0, dy = fabsf(y1 - y0), error = dx / 2.0f;
    int32_t x = (int32_t)x0, y = (int32_t)y0, end_x = (int32_t)x1, y_step = y0 < y1 ? 1 : -1;
    while (x <= end_x) {
      int32_t px = steep ? y : x, py = steep ? x : y;
      if (px >= 0 && px < mw && py >= 0 && py < mh) mask[(int64_t)py * mw + px] = value;
      x += 1;
      error -= dy;
      if (error < 0.0f) { y += y_step; error += dx; }
    }
  }
}

/* imageproc 0.22.0 draw_polygon_mut on a W x H Luma8 canvas (image_ops.rs:265-270) */
void orc_draw_polygon(uint8_t *canvas, int W, int H, const ipt *poly, int n, int value) {
  draw_polygon_mask(canvas, W, H, poly, n, (uint8_t)value);
}

double orc_box_score(const float *pred, int dim_m2, int dim_m1, const ipt *pts, int n,
                     int64_t *mask_count) {
  int64_t w = dim_m2, h = dim_m1; /* reference's (swapped) names */
  int64_t min_x = UINT32_MAX, max_x = 0, min_y = UINT32_MAX, max_y = 0;
  for (int i = 0; i < n; ++i) {
    if (pts[i].x < min_x) min_x = pts[i].x;
    if (pts[i].x > max_x) max_x = pts[i].x;
    if (pts[i].y < min_y) min_y = pts[i].y;
    if (pts[i].y > max_y) max_y = pts[i].y;
  }
  min_x = clampi(min_x, 0, w - 1); max_x = clampi(max_x, 0, w - 1);
  min_y = clampi(min_y, 0, h - 1); max_y = clampi(max_y, 0, h - 1);
  int mw = (int)(max_x - min_x + 1), mh = (int)(max_y - min_y + 1);
  uint8_t *mask = (uint8_t *)calloc((size_t)mw * mh, 1);
  ipt *moved = (ipt *)malloc(sizeof(ipt) * (size_t)n);
  for (int i = 0; i < n; ++i) { moved[i].x = pts[i].x - (int32_t)min_x; moved[i].y = pts[i].y - (int32_t)min_y; }
  draw_polygon_mask(mask, mw, mh, moved, n, 1);
  double s = 0.0, cnt = 0.0;
  for (int y = 0; y < mh; ++y)
    for (int x = 0; x < mw; ++x)
      if (mask[(int64_t)y * mw + x]) {
        float prod = pred[(int64_t)(min_y + y) * dim_m1 + (min_x + x)] * 1.0f;
        s += (double)prod;
        cnt += 1.0;
      }
  free(mask); free(moved);
  if (mask_count) *mask_count = (int64_t)cnt;
  return s / cnt;
}

/* ------------------------------------------------------------------------------------
 * ClipperOffset (Clipper 6.4.x) as driven by geo-clipper `offset(d, Miter(2.),
 * ClosedPolygon, 1.)` from polygon.rs:27-31.
 * ---------------------------------------------------------------------------------- */
static inline int64_t clip_round(double v) { return v < 0 ? (int64_t)(v - 0.5) : (int64_t)(v + 0.5); }

static double clipper_area(const ipt *p, int n) {
  if (n < 3) return 0;
  double a = 0;
  for (int i = 0, j = n - 1; i < n; ++i) {
    a += ((double)p[j].x + (double)p[i].x) * ((double)p[j].y - (double)p[i].y);
    j = i;
  }
  return -a * 0.5;
}

/* raw offset path (before the union clean-up). returns count; out capacity >= 3n */
static int clipper_offset_raw(const ipt *src_in, int n_in, double delta, ipt *out) {
  /* ClipperOffset::AddPath: strip closing / consecutive duplicates */
  ipt *src = (ipt *)malloc(sizeof(ipt) * (size_t)(n_in > 0 ? n_in : 1));
  int hi = n_in - 1;
  while (hi > 0 && src_in[0].x == src_in[hi].x && src_in[0].y == src_in[hi].y) hi--;
  int n = 0;
  for (int i = 0; i <= hi; ++i)
    if (n == 0 || src[n - 1].x != src_in[i].x || src[n - 1].y != src_in[i].y) src[n++] = src_in[i];
  if (n < 3) { free(src); return 0; }
  /* FixOrientations: reverse when Area < 0 */
  if (!(clipper_area(src, n) >= 0)) {
    for (int i = 0, j = n - 1; i < j; ++i, --j) { ipt t = src[i]; src[i] = src[j]; src[j] = t; }
  }
  int m = 0;
  if (fabs(delta) < 1.0e-20) { /* NEAR_ZERO: path copied unchanged */
    for (int i = 0; i < n; ++i) out[m++] = src[i];
    free(src);
    return m;
  }
  const double miter_lim = 0.5; /* MiterLimit 2 -> 2/(2*2) */
  double *nx = (double *)malloc(sizeof(double) * (size_t)n), *ny = (double *)malloc(sizeof(double) * (size_t)n);
  for (int j = 0; j < n; ++j) {
    ipt p1 = src[j], p2 = src[(j + 1) % n];
    if (p1.x == p2.x && p1.y == p2.y) { nx[j] = 0; ny[j] = 0; continue; }
    double dx = (double)(p2.x - p1.x), dy = (double)(p2.y - p1.y);
    double f = 1.0 / sqrt(dx * dx + dy * dy);
    dx *= f; dy *= f;
    nx[j] = dy; ny[j] = -dx;
  }
  int k = n - 1;
  for (int j = 0; j < n; ++j) {
    double sinA = nx[k] * ny[j] - nx[j] * ny[k];
    int done = 0;
    if (fabs(sinA * delta) < 1.0) {
      double cosA = nx[k] * nx[j] + ny[j] * ny[k];
      if (cosA > 0) {
        out[m].x = (int32_t)clip_round(src[j].x + nx[k] * delta);
        out[m].y = (int32_t)clip_round(src[j].y + ny[k] * delta);
        m++; done = 1;
      }
    } else if (sinA > 1.0) sinA = 1.0;
    else if (sinA < -1.0) sinA = -1.0;
    if (!done) {
      if (sinA * delta < 0) {
        out[m].x = (int32_t)clip_round(src[j].x + nx[k] * delta);
        out[m].y = (int32_t)clip_round(src[j].y + ny[k] * delta); m++;
        out[m++] = src[j];
        out[m].x = (int32_t)clip_round(src[j].x + nx[j] * delta);
        out[m].y = (int32_t)clip_round(src[j].y + ny[j] * delta); m++;
      } else {
        double r = 1 + (nx[j] * nx[k] + ny[j] * ny[k]);
        if (r >= miter_lim) { /* DoMiter */
          double q = delta / r;
          out[m].x = (int32_t)clip_round(src[j].x + (nx[k] + nx[j]) * q);
          out[m].y = (int32_t)clip_round(src[j].y + (ny[k] + ny[j]) * q); m++;
        } else { /* DoSquare */
          double dx = tan(atan2(sinA, nx[k] * nx[j] + ny[k] * ny[j]) / 4);
          out[m].x = (int32_t)clip_round(src[j].x + delta * (nx[k] - ny[k] * dx));
          out[m].y = (int32_t)clip_round(src[j].y + delta * (ny[k] + nx[k] * dx)); m++;
          out[m].x = (int32_t)clip_round(src[j].x + delta * (nx[j] + ny[j] * dx));
          out[m].y = (int32_t)clip_round(src[j].y + delta * (ny[j] - nx[j] * dx)); m++;
        }
      }
    }
    k = j;
  }
  free(nx); free(ny); free(src);
  return m;
}

/* Clipper IntersectPoint (main branch; scan-beam clamps not modelled). */
static void clipper_intersect_point(ipt a0, ipt a1, ipt b0, ipt b1, ipt *ip) {
  /* Bot = end with larger Y, Top = other; Dx = (Top.X-Bot.X)/(Top.Y-Bot.Y), HORIZONTAL=-1e40 */
  ipt abot, atop, bbot, btop;
  if (a0.y >= a1.y) { abot = a0; atop = a1; } else { abot = a1; atop = a0; }
  if (b0.y >= b1.y) { bbot = b0; btop = b1; } else { bbot = b1; btop = b0; }
  const double HORIZ = -1.0E+40;
  double adx = (atop.y == abot.y) ? HORIZ : (double)(atop.x - abot.x) / (double)(atop.y - abot.y);
  double bdx = (btop.y == bbot.y) ? HORIZ : (double)(btop.x - bbot.x) / (double)(btop.y - bbot.y);
  double b1_, b2_;
  int64_t X, Y;
  if (adx == bdx) { Y = abot.y; X = abot.x; }
  else if (adx == 0) {
    X = abot.x;
    if (bdx == HORIZ) Y = bbot.y;
    else { b2_ = bbot.y - (bbot.x / bdx); Y = clip_round(X / bdx + b2_); }
  } else if (bdx == 0) {
    X = bbot.x;
    if (adx == HORIZ) Y = abot.y;
    else { b1_ = abot.y - (abot.x / adx); Y = clip_round(X / adx + b1_); }
  } else {
    b1_ = abot.x - abot.y * adx;
    b2_ = bbot.x - bbot.y * bdx;
    double q = (b2_ - b1_) / (adx - bdx);
    Y = clip_round(q);
    if (fabs(adx) < fabs(bdx)) X = clip_round(adx * q + b1_);
    else X = clip_round(bdx * q + b2_);
  }
  ip->x = (int32_t)X; ip->y = (int32_t)Y;
}

typedef struct { int64_t num, den; } rat; /* den > 0 */
static inline int rat_lt(rat a, rat b) { return (__int128)a.num * b.den < (__int128)b.num * a.den; }
static inline int rat_eq(rat a, rat b) { return (__int128)a.num * b.den == (__int128)b.num * a.den; }
static inline int64_t crossi(int64_t ax, int64_t ay, int64_t bx, int64_t by) { return ax * by - ay * bx; }
static inline int64_t doti(int64_t ax, int64_t ay, int64_t bx, int64_t by) { return ax * bx + ay * by; }

/* Does segment j touch/cross segment i?  If so: t = parameter on i, s = parameter on j.
 * Collinear overlaps are reported only through j's END POINTS lying on i (which = 0/1).  */
static int seg_hit(const ipt *Q, int m, int i, int j, int which, rat *t, rat *s) {
  ipt a0 = Q[i], a1 = Q[(i + 1) % m], b0 = Q[j], b1 = Q[(j + 1) % m];
  int64_t dix = a1.x - a0.x, diy = a1.y - a0.y, djx = b1.x - b0.x, djy = b1.y - b0.y;
  int64_t wx = b0.x - a0.x, wy = b0.y - a0.y;
  int64_t den = crossi(dix, diy, djx, djy);
  if (den != 0) {
    if (which != 0) return 0;
    int64_t tn = crossi(wx, wy, djx, djy), sn = crossi(wx, wy, dix, diy);
    if (den < 0) { den = -den; tn = -tn; sn = -sn; }
    if (tn < 0 || tn > den || sn < 0 || sn > den) return 0;
    t->num = tn; t->den = den; s->num = sn; s->den = den;
    return 1;
  }
  if (crossi(wx, wy, dix, diy) != 0) return 0; /* parallel, not collinear */
  int64_t L = doti(dix, diy, dix, diy);
  if (L == 0) return 0;
  /* collinear: report end point `which` (1 -> b0 at s=0, 2 -> b1 at s=1) when it lies on i */
  if (which == 0) return 0;
  ipt e = which == 1 ? b0 : b1;
  int64_t tn = doti(e.x - a0.x, e.y - a0.y, dix, diy);
  if (tn < 0 || tn > L) return 0;
  t->num = tn; t->den = L; s->num = which == 1 ? 0 : 1; s->den = 1;
  return 1;
}

/* rotation rank of direction d relative to a reference direction r, counter-clockwise, in
 * (0, 2pi]: compares two candidate directions exactly.  returns 1 if a comes before b */
static int half_of(int64_t rx, int64_t ry, int64_t dx, int64_t dy) {
  /* 0: angle in (0,pi) ccw from r, 1: angle == pi, 2: (pi,2pi), 3: angle == 2pi (same as r) */
  int64_t c = crossi(rx, ry, dx, dy), d = doti(rx, ry, dx, dy);
  if (c > 0) return 0;
  if (c < 0) return 2;
  return d < 0 ? 1 : 3;
}
static int ccw_before(int64_t rx, int64_t ry, int64_t ax, int64_t ay, int64_t bx, int64_t by) {
  int ha = half_of(rx, ry, ax, ay), hb = half_of(rx, ry, bx, by);
  if (ha != hb) return ha < hb;
  if (ha == 1 || ha == 3) return 0;
  return crossi(ax, ay, bx, by) > 0; /* a before b when b is further ccw */
}
static int same_dir(int64_t ax, int64_t ay, int64_t bx, int64_t by) {
  return crossi(ax, ay, bx, by) == 0 && doti(ax, ay, bx, by) > 0;
}

/* One ray of the arrangement at a node: the part of path segment `seg` leaving the node
 * (sign +1, direction = the segment's) or arriving at it (sign -1, direction reversed).
 * Crossing a ray while turning counter-clockwise about the node changes the winding number
 * by its sign. */
typedef struct { int64_t dx, dy; int sign, seg; rat s; } ray_t;
#define ORC_MAX_RAYS 32

/* All rays at the point P = (pxn, pyn) / pden (pden > 0).  *is_vertex / *vtx: P coincides with
 * a path vertex.  Returns the ray count, -1 when more than ORC_MAX_RAYS meet in one point. */
static int rays_at(const ipt *Q, int m, __int128 pxn, __int128 pyn, __int128 pden, ray_t *rays,
                   int *is_vertex, ipt *vtx) {
  int k = 0;
  *is_vertex = 0;
  for (int j = 0; j < m; ++j) {
    ipt b0 = Q[j], b1 = Q[(j + 1) % m];
    int64_t dx = (int64_t)b1.x - b0.x, dy = (int64_t)b1.y - b0.y;
    __int128 qx = pxn - (__int128)b0.x * pden, qy = pyn - (__int128)b0.y * pden; /* (P - b0) * pden */
    if (qx * dy - qy * dx != 0) continue;                                        /* not on the line */
    /* parameter of P along the segment, from its dominant coordinate (fits 64 bits for 16-bit coordinates) */
    int use_x = (dx < 0 ? -dx : dx) >= (dy < 0 ? -dy : dy);
    int64_t dd = use_x ? dx : dy;
    __int128 sn = use_x ? qx : qy;
    if (dd < 0) sn = -sn;
    __int128 sd = pden * (dd < 0 ? -dd : dd);
    if (sn < 0 || sn > sd) continue;
    if (sn == 0) { *is_vertex = 1; *vtx = b0; }
    if (sn == sd) { *is_vertex = 1; *vtx = b1; }
    rat s = {(int64_t)sn, (int64_t)sd};
    if (sn < sd) { if (k >= ORC_MAX_RAYS) return -1; rays[k].dx = dx; rays[k].dy = dy; rays[k].sign = 1; rays[k].seg = j; rays[k].s = s; k++; }
    if (sn > 0) { if (k >= ORC_MAX_RAYS) return -1; rays[k].dx = -dx; rays[k].dy = -dy; rays[k].sign = -1; rays[k].seg = j; rays[k].s = s; k++; }
  }
  return k;
}

/* Turn counter-clockwise about a node, starting just after direction (rx, ry) where the
 * winding number is w0, and stop at the first group of coincident rays across which it
 * becomes positive: that group carries the boundary of {winding > 0} away from the node,
 * region on its left.  The group in direction r itself is met last.  Returns the index of an
 * outgoing ray of that group (-1: none) and the winding number on its right in *w_right. */
static int next_boundary_ray(ray_t *rays, int k, int64_t rx, int64_t ry, int w0, int prefer_seg, int *w_right) {
  for (int i = 1; i < k; ++i) { /* insertion sort, counter-clockwise from r */
    ray_t key = rays[i];
    int j = i - 1;
    while (j >= 0 && ccw_before(rx, ry, key.dx, key.dy, rays[j].dx, rays[j].dy)) { rays[j + 1] = rays[j]; j--; }
    rays[j + 1] = key;
  }
  int w = w0;
  for (int i = 0; i < k;) {
    int e = i, net = 0, pick = -1;
    while (e < k && same_dir(rays[i].dx, rays[i].dy, rays[e].dx, rays[e].dy)) {
      net += rays[e].sign;
      if (rays[e].sign > 0 && (pick < 0 || rays[e].seg == prefer_seg)) pick = e;
      e++;
    }
    if (w <= 0 && w + net > 0) { *w_right = w; return pick; }
    w += net;
    i = e;
  }
  return -1;
}

/* winding number of the closed path at (P.x - eps, P.y + delta), 0 < eps << delta << 1, for the
 * rational point P = (pxn, pyn) / pden: crossings of the upward vertical ray; segments through P
 * itself pass below that point */
static int winding_above_left(const ipt *Q, int m, __int128 pxn, __int128 pyn, __int128 pden) {
  int w = 0;
  for (int j = 0; j < m; ++j) {
    ipt a = Q[j], b = Q[(j + 1) % m];
    __int128 ax = (__int128)a.x * pden, bx = (__int128)b.x * pden;
    int dir;
    if (ax < pxn && pxn <= bx) dir = -1;      /* heading +x */
    else if (bx < pxn && pxn <= ax) dir = 1;  /* heading -x */
    else continue;
    /* y on the segment at x = P.x, compared with P.y: (P.x-a.x)*(b.y-a.y)/(b.x-a.x) > P.y-a.y */
    __int128 lhs = (pxn - ax) * ((int64_t)b.y - a.y), rhs = (pyn - (__int128)a.y * pden) * ((int64_t)b.x - a.x);
    int above = (b.x > a.x) ? (lhs > rhs) : (lhs < rhs);
    if (above) w += dir;
  }
  return w;
}

/* Is the node P on the boundary of {winding > 0}?  If so: the boundary ray leaving it. */
static int boundary_start_at(const ipt *Q, int m, __int128 pxn, __int128 pyn, __int128 pden, ray_t *rays,
                             int *seg, rat *s, int *w_right) {
  int isv; ipt vt;
  int k = rays_at(Q, m, pxn, pyn, pden, rays, &isv, &vt);
  if (k <= 0) return k;
  int w0 = winding_above_left(Q, m, pxn, pyn, pden);
  int pick = next_boundary_ray(rays, k, 0, 1, w0, -1, w_right); /* any w0: a rise is only seen after the winding was <= 0 */
  if (pick < 0) return 0;
  *seg = rays[pick].seg; *s = rays[pick].s;
  return 1;
}

/* node ordering of the start search: y descending, then x ascending; rational points */
typedef struct { __int128 xn, yn, den; } rpt;
static int node_after(rpt a, rpt b) { /* 1 if a comes strictly after b */
  __int128 ya = a.yn * b.den, yb = b.yn * a.den;
  if (ya != yb) return ya < yb;
  return a.xn * b.den > b.xn * a.den;
}

/*
 * Clipper's clean-up of ONE closed offset path = boundary of the region {winding > 0}
 * (ctUnion with pftPositive for delta > 0; for delta < 0 ClipperOffset adds a reversed outer
 * rectangle, fills pftNegative and returns the HOLES of that, which is the same region).
 * Winding numbers are taken in the raw (x, y) plane, counter-clockwise positive, so a path
 * with Clipper Area >= 0 has winding +1 inside.
 *
 * The region's boundary is walked with the region on the left.  At every node of the
 * arrangement (path vertices as they are, crossings as exact rationals) all rays through the
 * node are ordered by angle and the winding numbers of the sectors between them follow from
 * the one on the right of the arriving edge; the walk leaves along the first ray across which
 * the winding turns positive (next_boundary_ray).  Emitted vertices: path vertices as they
 * are, crossings through Clipper's IntersectPoint (double arithmetic, half-away-from-zero
 * rounding).  Afterwards duplicate and collinear vertices are dropped (FixupOutPolygon) and
 * the ring is rotated so that it starts right after the last top-most (min y, then max x)
 * vertex — Clipper's BuildResult order as observed on all golden polygons.
 *
 * Several polygons: Clipper sweeps from the largest y downwards and numbers output polygons
 * in the order their first (largest-y) vertex is met, merged pieces keeping the lower index;
 * the reference takes polygon 0 (polygon.rs:35), i.e. the one that owns the largest-y
 * boundary vertex (ties: smallest x).  The walk therefore starts at the first path vertex in
 * that order that lies on the region's boundary.  Holes are never visited (the reference
 * reads the exterior only).
 * Returns vertex count (0 = empty result).  out capacity >= 4*m+16.
 */
/* Walks one ring of the boundary from (start_seg, start_t) with w_right on its right, then applies
 * FixupOutPolygon.  Returns the vertex count (< 3: the ring collapsed on the integer grid), -1 on failure. */
static int walk_ring(const ipt *Q, int m, int start_seg, rat start_t, int w_right, ipt *out, int cap) {
  ray_t rays[ORC_MAX_RAYS];
  int cur = start_seg, n_out = 0;
  rat cur_t = start_t;
  int guard = 0, max_iter = 8 * m + 64;
  for (;;) {
    if (++guard > max_iter) return -1;
    /* next event on cur after cur_t */
    rat t_best = {1, 1};
    for (int j = 0; j < m; ++j) {
      if (j == cur) continue;
      for (int which = 0; which < 3; ++which) {
        rat t, s;
        if (!seg_hit(Q, m, cur, j, which, &t, &s)) continue;
        if (!rat_lt(cur_t, t)) continue;
        if (rat_lt(t, t_best)) t_best = t;
      }
    }
    ipt c0 = Q[cur], c1 = Q[(cur + 1) % m];
    int64_t ux = (int64_t)c1.x - c0.x, uy = (int64_t)c1.y - c0.y;
    /* node P = c0 + t_best * u */
    __int128 pden = t_best.den;
    __int128 pxn = (__int128)c0.x * pden + (__int128)t_best.num * ux, pyn = (__int128)c0.y * pden + (__int128)t_best.num * uy;
    int node_is_vertex; ipt node_v = {0, 0};
    int k = rays_at(Q, m, pxn, pyn, pden, rays, &node_is_vertex, &node_v);
    if (k < 0) return -1;
    int wr;
    int pick = next_boundary_ray(rays, k, -ux, -uy, w_right, cur, &wr);
    if (pick < 0) return -1;
    int nxt = rays[pick].seg;
    rat nxt_s = rays[pick].s;
    if (node_is_vertex || nxt != cur) {
      ipt node;
      if (node_is_vertex) node = node_v;
      else clipper_intersect_point(c0, c1, Q[nxt], Q[(nxt + 1) % m], &node);
      if (n_out >= cap) return -1;
      out[n_out++] = node;
    }
    if (nxt == start_seg && rat_eq(nxt_s, start_t)) break; /* closed the ring (the start node was emitted last) */
    cur = nxt; cur_t = nxt_s; w_right = wr;
  }
  /* FixupOutPolygon: drop duplicates and collinear middles until stable */
  int changed = 1;
  while (changed && n_out >= 3) {
    changed = 0;
    for (int i = 0; i < n_out && n_out >= 3; ++i) {
      ipt p = out[(i + n_out - 1) % n_out], c = out[i], n = out[(i + 1) % n_out];
      int dup = (c.x == n.x && c.y == n.y) || (c.x == p.x && c.y == p.y);
      int col = crossi((int64_t)c.x - p.x, (int64_t)c.y - p.y, (int64_t)n.x - c.x, (int64_t)n.y - c.y) == 0;
      if (dup || col) {
        memmove(out + i, out + i + 1, sizeof(ipt) * (size_t)(n_out - i - 1));
        n_out--; changed = 1; i--;
      }
    }
  }
  return n_out;
}

static int orc_union_positive(const ipt *Qin, int m_in, ipt *out, int cap) {
  ipt *Q = (ipt *)malloc(sizeof(ipt) * (size_t)(m_in > 0 ? m_in : 1));
  int m = 0;
  for (int i = 0; i < m_in; ++i)
    if (m == 0 || Q[m - 1].x != Qin[i].x || Q[m - 1].y != Qin[i].y) Q[m++] = Qin[i];
  while (m > 1 && Q[0].x == Q[m - 1].x && Q[0].y == Q[m - 1].y) m--;
  if (m < 3) { free(Q); return 0; }
  ray_t rays[ORC_MAX_RAYS];
  /* start search: nodes in (y descending, x ascending) order.  The first one is always a path
   * vertex (a crossing lies inside both segments' extents); the general search over vertices
   * and crossings runs only when that vertex is not on the region's boundary, or when the ring
   * through it collapses on the integer grid (Clipper disposes of output rings with fewer than
   * three distinct vertices, so "polygon 0" is the first SURVIVING one). */
  rpt last = {0, 0, 1};
  int n_out = 0;
  for (int tries = 0; tries < 4 * m + 16; ++tries) {
    rpt best = {0, 0, 0}; /* den == 0: none yet */
    for (int i = 0; i < m; ++i) {
      rpt c = {Q[i].x, Q[i].y, 1};
      if (tries > 0 && !node_after(c, last)) continue;
      if (best.den == 0 || node_after(best, c)) best = c;
    }
    if (tries > 0) {
      for (int i = 0; i < m; ++i)
        for (int j = i + 1; j < m; ++j) {
          rat t, sj;
          if (!seg_hit(Q, m, i, j, 0, &t, &sj)) continue;
          int64_t ux = (int64_t)Q[(i + 1) % m].x - Q[i].x, uy = (int64_t)Q[(i + 1) % m].y - Q[i].y;
          rpt c = {(__int128)Q[i].x * t.den + (__int128)t.num * ux, (__int128)Q[i].y * t.den + (__int128)t.num * uy, t.den};
          if (!node_after(c, last)) continue;
          if (best.den == 0 || node_after(best, c)) best = c;
        }
    }
    if (best.den == 0) break;
    last = best;
    int seg, w_right;
    rat st;
    int rc = boundary_start_at(Q, m, best.xn, best.yn, best.den, rays, &seg, &st, &w_right);
    if (rc < 0) break;
    if (rc == 0) continue;
    n_out = walk_ring(Q, m, seg, st, w_right, out, cap);
    if (n_out < 0) { n_out = 0; break; }
    if (n_out >= 3) break;
    n_out = 0;
  }
  free(Q);
  if (n_out < 3) return 0;
  /* BuildResult order: start right after the last top-most vertex */
  int top = 0;
  for (int i = 1; i < n_out; ++i)
    if (out[i].y < out[top].y || (out[i].y == out[top].y && out[i].x > out[top].x)) top = i;
  int st = (top + 1) % n_out;
  ipt *tmp = (ipt *)malloc(sizeof(ipt) * (size_t)n_out);
  for (int i = 0; i < n_out; ++i) tmp[i] = out[(st + i) % n_out];
  memcpy(out, tmp, sizeof(ipt) * (size_t)n_out);
  free(tmp);
  return n_out;
}

/* raw offset path -> cleaned polygon (test hook for the independent region check) */
int orc_union_of_path(const ipt *raw, int m, ipt *out, int cap) { return m >= 3 ? orc_union_positive(raw, m, out, cap) : 0; }

/* clip_polygon(points, factor, Shrink | Expand) — polygon.rs:13-49.  Returns vertex count, 0 = None. */
int orc_clip_polygon(const ipt *pts, int n, double factor, int shrink, ipt *out, int cap, double *distance_out) {
  /* geo 0.15: unsigned_area (shoelace, closed ring) and euclidean_length of the closed ring */
  double twice = 0.0, perim = 0.0;
  for (int i = 0; i < n; ++i) {
    ipt a = pts[i], b = pts[(i + 1) % n];
    twice += (double)a.x * (double)b.y - (double)a.y * (double)b.x;
    perim += pt_dist(a, b);
  }
  double area = fabs(twice / 2.0);
  double distance = area * factor / perim;
  if (shrink) distance *= -1.;
  if (distance_out) *distance_out = distance;
  ipt *raw = (ipt *)malloc(sizeof(ipt) * (size_t)(3 * n + 3));
  int m = clipper_offset_raw(pts, n, distance, raw);
  int r = m >= 3 ? orc_union_positive(raw, m, out, cap) : 0;
  free(raw);
  return r;
}

/* expand_polygon(points, factor) — polygon.rs:51-56 */
int orc_expand_polygon(const ipt *pts, int n, double factor, ipt *out, int cap, double *distance_out) {
  return orc_clip_polygon(pts, n, factor, 0, out, cap, distance_out);
}

int orc_offset_raw(const ipt *pts, int n, double delta, ipt *out) { return clipper_offset_raw(pts, n, delta, out); }

/* ------------------------------------------------------------------------------------
 * min_area_rect + get_min_area_bounding_box — imageproc geometry.rs / metrics.rs:133-148
 * (SURVEY A.5).  Returns the short side; box[4] receives the ordered corners.
 * ---------------------------------------------------------------------------------- */
typedef struct { double x, y; } dpt;

static int orient(dpt p, dpt q, dpt r) {
  double val = (q.y - p.y) * (r.x - q.x) - (q.x - p.x) * (r.y - q.y);
  if (val == 0.0) return 0;
  return val > 0.0 ? 1 : 2; /* 1 clockwise, 2 counter-clockwise */
}
static double ddist(dpt a, dpt b) { return sqrt((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y)); }

static dpt g_anchor;
static int hull_cmp(const void *pa, const void *pb) {
  dpt a = *(const dpt *)pa, b = *(const dpt *)pb;
  int o = orient(g_anchor, a, b);
  if (o == 0) return ddist(g_anchor, a) < ddist(g_anchor, b) ? -1 : 1;
  return o == 2 ? -1 : 1;
}

static int convex_hull(const ipt *pts, int n, dpt *hull) {
  if (n == 0) return 0;
  dpt *p = (dpt *)malloc(sizeof(dpt) * (size_t)n);
  for (int i = 0; i < n; ++i) { p[i].x = pts[i].x; p[i].y = pts[i].y; }
  int s = 0;
  for (int i = 1; i < n; ++i)
    if (p[i].y < p[s].y || (p[i].y == p[s].y && p[i].x < p[s].x)) s = i;
  dpt start = p[s];
  p[s] = p[0]; /* points.swap(0, pos); points.remove(0) */
  dpt *rest = p + 1;
  int nr = n - 1;
  g_anchor = start;
  /* stable merge sort is what Rust's sort_by is; emulate stability with insertion sort */
  for (int i = 1; i < nr; ++i) {
    dpt key = rest[i];
    int j = i - 1;
    while (j >= 0 && hull_cmp(&key, &rest[j]) < 0) { rest[j + 1] = rest[j]; j--; }
    rest[j + 1] = key;
  }
  /* drop collinear (keep farthest) */
  dpt *rem = (dpt *)malloc(sizeof(dpt) * (size_t)(nr > 0 ? nr : 1));
  int nrem = 0;
  for (int i = 0; i < nr;) {
    int k = i;
    while (k + 1 < nr && orient(start, rest[k], rest[k + 1]) == 0) k++;
    rem[nrem++] = rest[k];
    i = k + 1;
  }
  int h = 0;
  hull[h++] = start;
  for (int i = 0; i < nrem; ++i) {
    while (h > 1 && orient(hull[h - 2], hull[h - 1], rem[i]) != 2) h--;
    hull[h++] = rem[i];
  }
  free(rem); free(p);
  return h;
}

static dpt rot(dpt p, double s, double c) { dpt r; r.x = p.x * c + p.y * s; r.y = p.y * c - p.x * s; return r; }
static dpt irot(dpt p, double s, double c) { dpt r; r.x = p.x * c - p.y * s; r.y = p.y * c + p.x * s; return r; }

static void min_area_rect(const ipt *pts, int n, ipt box[4]) {
  dpt *hull = (dpt *)malloc(sizeof(dpt) * (size_t)(n + 1));
  int h = convex_hull(pts, n, hull);
  if (h == 1) { for (int i = 0; i < 4; ++i) { box[i].x = (int32_t)hull[0].x; box[i].y = (int32_t)hull[0].y; } free(hull); return; }
  if (h == 2) {
    box[0].x = (int32_t)hull[0].x; box[0].y = (int32_t)hull[0].y;
    box[1].x = (int32_t)hull[1].x; box[1].y = (int32_t)hull[1].y;
    box[2] = box[1]; box[3] = box[0]; free(hull); return;
  }
  const double PI = 3.14159265358979323846264338327950288;
  double min_area = 1.7976931348623157e308;
  dpt res[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  for (int e = 0; e + 1 < h; ++e) { /* windows(2): no wrap-around edge */
    double ex = hull[e + 1].x - hull[e].x, ey = hull[e + 1].y - hull[e].y;
    double angle = fabs(fmod(atan2(ey, ex) + PI, PI / 2.));
    double s = sin(angle), c = cos(angle);
    double min_x = 1.7976931348623157e308, max_x = -1.7976931348623157e308;
    double min_y = 1.7976931348623157e308, max_y = -1.7976931348623157e308;
    for (int i = 0; i < h; ++i) {
      dpt r = rot(hull[i], s, c);
      if (r.x < min_x) min_x = r.x;
      if (r.x > max_x) max_x = r.x;
      if (r.y < min_y) min_y = r.y;
      if (r.y > max_y) max_y = r.y;
    }
    double area = (max_x - min_x) * (max_y - min_y);
    if (area < min_area) {
      min_area = area;
      dpt a = {max_x, min_y}, b = {min_x, min_y}, cc = {min_x, max_y}, d = {max_x, max_y};
      res[0] = irot(a, s, c); res[1] = irot(b, s, c); res[2] = irot(cc, s, c); res[3] = irot(d, s, c);
    }
  }
  /* stable sort by x */
  for (int i = 1; i < 4; ++i) {
    dpt key = res[i]; int j = i - 1;
    while (j >= 0 && key.x < res[j].x) { res[j + 1] = res[j]; j--; }
    res[j + 1] = key;
  }
  int i1 = res[1].y > res[0].y ? 0 : 1;
  int i2 = res[3].y > res[2].y ? 2 : 3;
  int i3 = res[3].y > res[2].y ? 3 : 2;
  int i4 = res[1].y > res[0].y ? 1 : 0;
  box[0].x = (int32_t)floor(res[i1].x); box[0].y = (int32_t)floor(res[i1].y);
  box[1].x = (int32_t)ceil(res[i2].x);  box[1].y = (int32_t)floor(res[i2].y);
  box[2].x = (int32_t)ceil(res[i3].x);  box[2].y = (int32_t)ceil(res[i3].y);
  box[3].x = (int32_t)floor(res[i4].x); box[3].y = (int32_t)ceil(res[i4].y);
  free(hull);
}

double orc_min_area_bounding_box(const ipt *pts, int n, ipt out_box[4]) {
  ipt b[4];
  min_area_rect(pts, n, b);
  /* metrics.rs:137-147: stable sort by x, reorder, short side */
  for (int i = 1; i < 4; ++i) {
    ipt key = b[i]; int j = i - 1;
    while (j >= 0 && key.x < b[j].x) { b[j + 1] = b[j]; j--; }
    b[j + 1] = key;
  }
  int i1 = b[1].y > b[0].y ? 0 : 1;
  int i2 = b[3].y > b[2].y ? 2 : 3;
  int i3 = b[3].y > b[2].y ? 3 : 2;
  int i4 = b[1].y > b[0].y ? 1 : 0;
  ipt r[4] = {b[i1], b[i2], b[i3], b[i4]};
  if (out_box) memcpy(out_box, r, sizeof(r));
  double w = pt_dist(r[0], r[1]), h = pt_dist(r[0], r[3]);
  return w < h ? w : h;
}

/* ------------------------------------------------------------------------------------
 * get_polygons_from_bitmap — metrics.rs:58-127, one image.
 * pred [H,W] f32, bitmap [H,W] u8 {0,1} (the reference multiplies by 255 first; only >0
 * matters).  Output polygons as u32 xy pairs, offsets[n+1] in points, scores[n].
 * stats[0..4] = contours, >=4 dp points, >= box_thresh, kept, dropped-empty-offset.
 * Returns polygon count or -1 on capacity overflow.
 * ---------------------------------------------------------------------------------- */
static inline uint32_t sat_u32(double v) {
  if (!(v == v)) return 0;
  if (v <= 0.0) return 0;
  if (v >= 4294967295.0) return 4294967295u;
  return (uint32_t)v;
}

int orc_polygons_from_bitmap(const float *pred, const uint8_t *bitmap, int H, int W,
                             double adj_x, double adj_y, double box_thresh, double min_size,
                             double unclip_factor, int max_polys, int64_t max_pts,
                             int64_t *offsets, uint32_t *xy, double *scores, int64_t *stats, ipt *boxes_out) {
  int64_t cap_pts = (int64_t)W * H * 2 + 16;
  int cap_c = W * H / 2 + 16;
  ipt *cpts = (ipt *)malloc(sizeof(ipt) * (size_t)cap_pts);
  int64_t *coff = (int64_t *)malloc(sizeof(int64_t) * (size_t)(cap_c + 1));
  uint8_t *ctype = (uint8_t *)malloc((size_t)cap_c);
  int nc = orc_find_contours(bitmap, W, H, cpts, cap_pts, coff, ctype, cap_c);
  int n_out = 0;
  int64_t np_out = 0;
  int64_t st[5] = {0, 0, 0, 0, 0};
  offsets[0] = 0;
  int rc = 0;
  if (nc < 0) { rc = -1; goto done; }
  st[0] = nc;
  for (int ci = 0; ci < nc; ++ci) {
    const ipt *c = cpts + coff[ci];
    int64_t len = coff[ci + 1] - coff[ci];
    double eps = 0.01 * orc_arc_length(c, len, 1);
    if (eps == 0.) eps = 0.01;
    ipt *dp = (ipt *)malloc(sizeof(ipt) * (size_t)(len + 2));
    int64_t nd = orc_approx_dp(c, len, eps, 1, dp);
    if (nd > 1 && dp[0].x == dp[nd - 1].x && dp[0].y == dp[nd - 1].y) nd--;
    if (nd < 4) { free(dp); continue; }
    st[1]++;
    double score = orc_box_score(pred, H, W, dp, (int)nd, NULL);
    if (box_thresh > score) { free(dp); continue; }
    st[2]++;
    int cap = (int)(6 * nd + 32); /* same bound as the CUDA path (geometry.cu unclip_cap) */
    ipt *ex = (ipt *)malloc(sizeof(ipt) * (size_t)cap);
    int ne = orc_expand_polygon(dp, (int)nd, unclip_factor, ex, cap, NULL);
    free(dp);
    if (ne == 0) { st[4]++; free(ex); continue; } /* reference: unwrap() panic (D11) */
    ipt box[4];
    double sside = orc_min_area_bounding_box(ex, ne, box);
    if (sside < min_size) { free(ex); continue; }
    if (n_out >= max_polys || np_out + ne > max_pts) { free(ex); rc = -1; goto done; }
    for (int i = 0; i < ne; ++i) {
      xy[2 * (np_out + i)] = sat_u32(round((double)ex[i].x / adj_x));
      xy[2 * (np_out + i) + 1] = sat_u32(round((double)ex[i].y / adj_y));
    }
    np_out += ne;
    scores[n_out] = score;
    if (boxes_out) memcpy(boxes_out + 4 * (size_t)n_out, box, sizeof(box)); /* map coordinates; input of the crop glue */
    n_out++;
    offsets[n_out] = np_out;
    st[3]++;
    free(ex);
  }
done:
  if (stats) memcpy(stats, st, sizeof(st));
  free(cpts); free(coff); free(ctype);
  return rc < 0 ? rc : n_out;
}

/* ------------------------------------------------------------------------------------
 * preprocess_image resize part — image 0.23.11 DynamicImage::resize(Triangle) + to_luma +
 * zero pad (image_ops.rs:188-220).  src RGBA8 [sh, sw, 4].
 * Writes dst [H, W] u8 and the resized dims; adjust = resized / original.
 * ---------------------------------------------------------------------------------- */
static inline float tri(float x) { float a = fabsf(x); return a < 1.0f ? 1.0f - a : 0.0f; }

void orc_resize_dims(int sw, int sh, int W, int H, int *rw, int *rh) {
  /* image 0.23.11 math::utils::resize_dimensions(width, height, nwidth, nheight, fill=false):
   * integer arithmetic */
  uint64_t ratio = (uint64_t)sw * (uint64_t)H, nratio = (uint64_t)W * (uint64_t)sh;
  int use_width = nratio <= ratio;
  uint64_t inter = use_width ? (uint64_t)sh * (uint64_t)W / (uint64_t)sw
                             : (uint64_t)sw * (uint64_t)H / (uint64_t)sh;
  if (inter < 1) inter = 1;
  if (use_width) { *rw = W; *rh = (int)inter; } else { *rw = (int)inter; *rh = H; }
}

int g_resize_norm_first = 0; /* image 0.23.11 divides by the weight sum after accumulation (unpinned within JPEG-decoder noise, see tests) */
void orc_set_resize_variant(int norm_first) { g_resize_norm_first = norm_first; }

static void resample_1d(const uint8_t *src, int n_in, int n_out, int stride_in, int stride_out,
                        int lines, int line_stride_in, int line_stride_out, int channels, uint8_t *dst) {
  float ratio = (float)n_in / (float)n_out;
  float sratio = ratio < 1.0f ? 1.0f : ratio;
  float support = 1.0f * sratio;
  for (int o = 0; o < n_out; ++o) {
    float inputc = ((float)o + 0.5f) * ratio;
    int64_t left = (int64_t)floorf(inputc - support);
    if (left < 0) left = 0; if (left > n_in - 1) left = n_in - 1;
    int64_t right = (int64_t)ceilf(inputc + support);
    if (right < left + 1) right = left + 1; if (right > n_in) right = n_in;
    float inputc2 = inputc - 0.5f;
    float ws[64]; /* support <= ~ratio+2; callers keep ratio small enough */
    int nw = (int)(right - left);
    float *wv = nw <= 64 ? ws : (float *)malloc(sizeof(float) * (size_t)nw);
    float sum = 0.0f;
    for (int i = 0; i < nw; ++i) { wv[i] = tri(((float)(left + i) - inputc2) / sratio); sum += wv[i]; }
    if (g_resize_norm_first) for (int i = 0; i < nw; ++i) wv[i] /= sum;
    for (int l = 0; l < lines; ++l) {
      for (int ch = 0; ch < channels; ++ch) {
        float t = 0.0f;
        for (int i = 0; i < nw; ++i)
          t += (float)src[(int64_t)l * line_stride_in + (left + i) * stride_in + ch] * wv[i];
        if (!g_resize_norm_first) t = t / sum;
        float cl = t < 0.0f ? 0.0f : (t > 255.0f ? 255.0f : t);
        dst[(int64_t)l * line_stride_out + (int64_t)o * stride_out + ch] = (uint8_t)cl; /* truncation */
      }
    }
    if (wv != ws) free(wv);
  }
}

void orc_preprocess(const uint8_t *rgba, int sw, int sh, int W, int H, uint8_t *dst,
                    int *rw_out, int *rh_out) {
  int rw, rh;
  orc_resize_dims(sw, sh, W, H, &rw, &rh);
  *rw_out = rw; *rh_out = rh;
  memset(dst, 0, (size_t)W * H);
  uint8_t *res;
  if (rw == sw && rh == sh) {
    res = (uint8_t *)malloc((size_t)sw * sh * 4);
    memcpy(res, rgba, (size_t)sw * sh * 4);
  } else {
    /* vertical pass first: [sh, sw] -> [rh, sw]; then horizontal: -> [rh, rw] */
    uint8_t *tmp = (uint8_t *)malloc((size_t)sw * rh * 4);
    resample_1d(rgba, sh, rh, sw * 4, sw * 4, sw, 4, 4, 4, tmp);
    res = (uint8_t *)malloc((size_t)rw * rh * 4);
    resample_1d(tmp, sw, rw, 4, 4, rh, sw * 4, rw * 4, 4, res);
    free(tmp);
  }
  for (int y = 0; y < rh && y < H; ++y)
    for (int x = 0; x < rw && x < W; ++x) {
      const uint8_t *p = res + ((int64_t)y * rw + x) * 4;
      float l = 0.2126f * (float)p[0] + 0.7152f * (float)p[1] + 0.0722f * (float)p[2];
      dst[(int64_t)y * W + x] = (uint8_t)l;
    }
  free(res);
}

/* ------------------------------------------------------------------------------------
 * Polygon -> glyph crop glue ("crop spec v1").  The reference has NO counterpart: character
 * segmentation is an open item of its README (README.md:20-26), and its recognition net is
 * only ever fed ready-made 28x28 files (image_ops.rs:73-85).  This is the definition the
 * CUDA path (csrc/crop.cu) is held to, built from the reference's own pieces: the box is
 * get_min_area_bounding_box's (metrics.rs:133-148; TL, TR, BR, BL), the resampling is
 * preprocess_image's Triangle filter (image 0.23.11).
 *
 *   1. reading axis = the longer side of the box (u = TR - TL, v = BL - TL; swapped when |v| > |u|,
 *      so vertical text reads top to bottom); pw = max(1, round(|u|)), ph = max(1, round(|v|)), f32
 *   2. the box is rectified to a pw x ph patch by nearest sampling:
 *      patch(x, y) = img[clamp(floor(o.y + s uy + t vy))][clamp(floor(o.x + s ux + t vx))],
 *      s = (x + 0.5) / pw, t = (y + 0.5) / ph, every operation a separately rounded f32 one
 *   3. the patch is cut into K cells along x: cell g = columns [min(g pw / K, pw - 1), max(x0 + 1, min(pw, (g + 1) pw / K)))
 *   4. each cell is resized to 28 x 28 with the Triangle filter, vertical pass first, u8 truncation
 *      after either pass (exactly resample_1d above)
 * out: [K][784] u8.
 * ---------------------------------------------------------------------------------- */
void orc_crop_glyphs(const uint8_t *img, int W, int H, const ipt box[4], int K, uint8_t *out) {
  ipt o = box[0];
  float ux = (float)(box[1].x - box[0].x), uy = (float)(box[1].y - box[0].y);
  float vx = (float)(box[3].x - box[0].x), vy = (float)(box[3].y - box[0].y);
  float wlen = sqrtf(ux * ux + uy * uy), hlen = sqrtf(vx * vx + vy * vy);
  if (hlen > wlen) {
    float t;
    t = ux; ux = vx; vx = t;
    t = uy; uy = vy; vy = t;
    t = wlen; wlen = hlen; hlen = t;
  }
  int pw = (int)roundf(wlen), ph = (int)roundf(hlen);
  if (pw < 1) pw = 1;
  if (ph < 1) ph = 1;
  uint8_t *patch = (uint8_t *)malloc((size_t)pw * ph);
  for (int y = 0; y < ph; ++y)
    for (int x = 0; x < pw; ++x) {
      float s = ((float)x + 0.5f) / (float)pw, t = ((float)y + 0.5f) / (float)ph;
      float fx = ((float)o.x + s * ux) + t * vx, fy = ((float)o.y + s * uy) + t * vy;
      int ix = (int)floorf(fx), iy = (int)floorf(fy);
      ix = ix < 0 ? 0 : (ix > W - 1 ? W - 1 : ix);
      iy = iy < 0 ? 0 : (iy > H - 1 ? H - 1 : iy);
      patch[(size_t)y * pw + x] = img[(size_t)iy * W + ix];
    }
  int saved = g_resize_norm_first;
  g_resize_norm_first = 0;
  for (int g = 0; g < K; ++g) {
    int x0 = (int)((int64_t)g * pw / K);
    if (x0 > pw - 1) x0 = pw - 1;
    int x1 = (int)((int64_t)(g + 1) * pw / K);
    if (x1 > pw) x1 = pw;
    if (x1 < x0 + 1) x1 = x0 + 1;
    int sw = x1 - x0;
    uint8_t *tmp = (uint8_t *)malloc((size_t)28 * sw);
    /* vertical: [ph][sw] (row stride pw) -> [28][sw] */
    resample_1d(patch + x0, ph, 28, pw, sw, sw, 1, 1, 1, tmp);
    /* horizontal: [28][sw] -> [28][28] */
    resample_1d(tmp, sw, 28, 1, 1, 28, sw, 28, 1, out + (size_t)g * 784);
    free(tmp);
  }
  g_resize_norm_first = saved;
  free(patch);
}
