// Per-candidate geometry of the detection post-processing, on the GPU:
//   box_score_fast            metrics.rs:150-184  (imageproc draw_polygon_mut mask, f64 mean)
//   expand_polygon            polygon.rs:13-56    (ClipperOffset miter join + union clean-up)
//   get_min_area_bounding_box metrics.rs:133-148  (imageproc min_area_rect)
//   rescale / round / emit    metrics.rs:109-123
// Compiled with -fmad=false: every f32/f64 expression must round like the reference's
// scalar code.  Integer predicates are exact (int64 / __int128).
#include <vector>

#include "common.cuh"
#include "dd_math.cuh"

namespace ocrb {

struct ipt { int x, y; };
struct dpt { double x, y; };

// =======================================================================================
// box score: one CTA per candidate.  The polygon mask (scan-line fill with f32-rounded
// crossings + Bresenham outline, union) is built as a bit mask in shared memory, in bands
// of rows when the bounding box is larger than the shared-memory budget; the masked f32
// probabilities are accumulated in f64 in a fixed order (deterministic).
// HBM traffic: bbox_area * 4 B of the probability map, read once.
// =======================================================================================
constexpr int BS_THREADS = 256;
constexpr int BS_MAX_PTS = 256;           // DP polygons with more vertices are read from global memory
constexpr int BS_MASK_WORDS = 8192;       // 32 KB of mask bits per band

__device__ __forceinline__ int clampi(long long v, long long lo, long long hi) { return (int)(v < lo ? lo : (v > hi ? hi : v)); }

__global__ void __launch_bounds__(BS_THREADS) box_score_kernel(const float *__restrict__ pred, int dim_m2, int dim_m1,
                                                               int64_t image_stride, const int *__restrict__ cand_contour,
                                                               const int64_t *__restrict__ start_idx,
                                                               const int64_t *__restrict__ chain_off,
                                                               const ushort2 *__restrict__ dp_pts, const int *__restrict__ dp_count,
                                                               int n_cand, double *__restrict__ scores, int *__restrict__ err_flags) {
  __shared__ uint32_t mask[BS_MASK_WORDS];
  __shared__ int px[BS_MAX_PTS], py[BS_MAX_PTS];
  __shared__ double red_s[BS_THREADS / 32];
  __shared__ long long red_c[BS_THREADS / 32];
  const int cand = blockIdx.x;
  if (cand >= n_cand) return;
  const int c = cand_contour ? cand_contour[cand] : cand;
  const int n = dp_count[c];
  if (n < 1) {
    if (threadIdx.x == 0) scores[cand] = -1.0;
    return;
  }
  const ushort2 *pts = dp_pts + chain_off[c];
  const int64_t b = start_idx ? start_idx[c] / image_stride : 0;
  const float *map = pred + b * image_stride;
  // bounding box with the reference's clamps (x by size[-2]-1, y by size[-1]-1; D10)
  long long mnx = 0xffffffffll, mxx = 0, mny = 0xffffffffll, mxy = 0;
  for (int i = 0; i < n; ++i) {
    int x = pts[i].x, y = pts[i].y;
    mnx = x < mnx ? x : mnx; mxx = x > mxx ? x : mxx;
    mny = y < mny ? y : mny; mxy = y > mxy ? y : mxy;
  }
  const int min_x = clampi(mnx, 0, dim_m2 - 1), max_x = clampi(mxx, 0, dim_m2 - 1);
  const int min_y = clampi(mny, 0, dim_m1 - 1), max_y = clampi(mxy, 0, dim_m1 - 1);
  const int mw = max_x - min_x + 1, mh = max_y - min_y + 1;
  // vertices relative to the box: in shared memory when they fit, straight from the DP arena otherwise (the
  // reference has no vertex limit; a polygon with more than BS_MAX_PTS vertices is just slower here)
  const bool in_smem = n <= BS_MAX_PTS;
  if (in_smem)
    for (int i = threadIdx.x; i < n; i += BS_THREADS) { px[i] = (int)pts[i].x - min_x; py[i] = (int)pts[i].y - min_y; }
  __syncthreads();
  auto PX = [&](int i) { return in_smem ? px[i] : (int)pts[i].x - min_x; };
  auto PY = [&](int i) { return in_smem ? py[i] : (int)pts[i].y - min_y; };
  // polygon vertical range clipped to the canvas (draw_polygon_mut)
  int y_min = INT32_MAX, y_max = INT32_MIN;
  for (int i = 0; i < n; ++i) { y_min = min(y_min, PY(i)); y_max = max(y_max, PY(i)); }
  y_min = max(0, min(y_min, mh - 1));
  y_max = max(0, min(y_max, mh - 1));

  const int row_words = (mw + 31) >> 5;
  int band_rows = BS_MASK_WORDS / row_words;
  if (band_rows < 1) {  // a single row does not fit (bbox wider than 262144 px): unsupported
    if (threadIdx.x == 0) { scores[cand] = -1.0; atomicOr(err_flags, 2); }
    return;
  }
  double acc = 0.0;
  long long cnt = 0;
  for (int yb0 = 0; yb0 < mh; yb0 += band_rows) {
    const int yb1 = min(mh, yb0 + band_rows);
    const int words = (yb1 - yb0) * row_words;
    for (int i = threadIdx.x; i < words; i += BS_THREADS) mask[i] = 0;
    __syncthreads();
    // ---- scan-line fill: one thread per row ----
    for (int y = max(yb0, y_min) + threadIdx.x; y <= min(yb1 - 1, y_max); y += BS_THREADS) {
      uint32_t *rowm = mask + (y - yb0) * row_words;
      // crossings in ascending order are consumed pairwise; selection by repeated minimum
      // keeps memory O(1): k-th smallest with multiplicity via (value, rank) stepping.
      // n is small (<= 256), so an O(n^2) pass per row is fine.
      int last_v = INT32_MIN, last_taken = 0;  // how many copies of last_v already consumed
      int span_from = 0;
      bool have_from = false;
      for (;;) {
        // find the next crossing value >= last_v (respecting multiplicities)
        int best = INT32_MAX, best_mult = 0, cur_mult = 0;
        for (int e = 0; e < n; ++e) {
          const int e1 = e + 1 == n ? 0 : e + 1;
          int x0 = PX(e), y0 = PY(e), x1 = PX(e1), y1 = PY(e1);
          if (!((y0 <= y && y1 >= y) || (y1 <= y && y0 >= y))) continue;
          int v[2], nv = 0;
          if (y0 == y1) { v[nv++] = x0; v[nv++] = x1; }
          else if (y0 == y || y1 == y) {
            if (y1 > y) v[nv++] = x0;
            if (y0 > y) v[nv++] = x1;
          } else {
            float fraction = (float)(y - y0) / (float)(y1 - y0);
            float inter = (float)x0 + fraction * (float)(x1 - x0);
            v[nv++] = (int)roundf(inter);
          }
          for (int q = 0; q < nv; ++q) {
            if (v[q] == last_v) cur_mult++;
            else if (v[q] > last_v) {
              if (v[q] < best) { best = v[q]; best_mult = 1; }
              else if (v[q] == best) best_mult++;
            }
          }
        }
        int val, avail;
        if (last_v != INT32_MIN && cur_mult > last_taken) { val = last_v; avail = cur_mult - last_taken; }
        else if (best != INT32_MAX) { val = best; avail = best_mult; last_v = best; last_taken = 0; }
        else break;
        // consume all `avail` copies of val
        for (int k = 0; k < avail; ++k) {
          if (!have_from) { span_from = val; have_from = true; }
          else {
            int from = min(span_from, mw), to = min(val, mw - 1);
            if (from < mw && to >= 0) {
              from = max(0, from); to = max(0, to);
              for (int x = from; x <= to; ++x) rowm[x >> 5] |= 1u << (x & 31);
            }
            have_from = false;
          }
        }
        last_taken += avail;
      }
    }
    __syncthreads();
    // ---- outline: one thread per edge, Bresenham exactly as imageproc (f32 state) ----
    for (int e = threadIdx.x; e < n; e += BS_THREADS) {
      const int e1 = e + 1 == n ? 0 : e + 1;
      float x0 = (float)PX(e), y0 = (float)PY(e);
      float x1 = (float)PX(e1), y1 = (float)PY(e1);
      bool steep = fabsf(y1 - y0) > fabsf(x1 - x0);
      if (steep) { float t = x0; x0 = y0; y0 = t; t = x1; x1 = y1; y1 = t; }
      if (x0 > x1) { float t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; }
      float dx = x1 - x0, dy = fabsf(y1 - y0), error = dx / 2.0f;
      int x = (int)x0, y = (int)y0, end_x = (int)x1, y_step = y0 < y1 ? 1 : -1;
      while (x <= end_x) {
        int qx = steep ? y : x, qy = steep ? x : y;
        if (qx >= 0 && qx < mw && qy >= yb0 && qy < yb1) atomicOr(&mask[(qy - yb0) * row_words + (qx >> 5)], 1u << (qx & 31));
        x += 1;
        error -= dy;
        if (error < 0.0f) { y += y_step; error += dx; }
      }
    }
    __syncthreads();
    // ---- masked sum ----
    for (int i = threadIdx.x; i < words; i += BS_THREADS) {
      uint32_t m = mask[i];
      if (!m) continue;
      int ry = i / row_words, wx = (i % row_words) << 5;
      const float *rowp = map + (int64_t)(min_y + yb0 + ry) * dim_m1 + (min_x + wx);
      cnt += __popc(m);
      while (m) {
        int bit = __ffs(m) - 1;
        m &= m - 1;
        acc += (double)rowp[bit];
      }
    }
    __syncthreads();
  }
  // deterministic block reduction
  for (int d = 16; d > 0; d >>= 1) {
    acc += __shfl_down_sync(0xffffffffu, acc, d);
    cnt += __shfl_down_sync(0xffffffffu, cnt, d);
  }
  if ((threadIdx.x & 31) == 0) { red_s[threadIdx.x >> 5] = acc; red_c[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    long long k = 0;
    for (int w = 0; w < BS_THREADS / 32; ++w) { s += red_s[w]; k += red_c[w]; }
    scores[cand] = s / (double)k;
  }
}

int launch_box_score(ocrb_ctx *ctx, const float *pred, int dim_m2, int dim_m1, int64_t image_stride,
                     const int *cand_contour, const int64_t *start_idx, const int64_t *chain_off, const ushort2 *dp_pts,
                     const int *dp_count, int n_cand, double *scores, int *err_flags) {
  if (n_cand <= 0) return OCRB_OK;
  box_score_kernel<<<n_cand, BS_THREADS, 0, ctx->stream>>>(pred, dim_m2, dim_m1, image_stride, cand_contour, start_idx,
                                                           chain_off, dp_pts, dp_count, n_cand, scores, err_flags);
  return check_launch(ctx, "box_score");
}

// =======================================================================================
// unclip: ClipperOffset (miter limit 2) + union/pftPositive clean-up + min-area-rect.
// One thread per candidate that passed the score filter; scratch slabs in global memory.
// =======================================================================================
__host__ __device__ __forceinline__ long long clip_round(double v) { return v < 0 ? (long long)(v - 0.5) : (long long)(v + 0.5); }
__host__ __device__ __forceinline__ long long crossi(long long ax, long long ay, long long bx, long long by) { return ax * by - ay * bx; }
__host__ __device__ __forceinline__ long long doti(long long ax, long long ay, long long bx, long long by) { return ax * bx + ay * by; }

__host__ __device__ double clipper_area(const ipt *p, int n) {
  if (n < 3) return 0;
  double a = 0;
  for (int i = 0, j = n - 1; i < n; ++i) {
    a += ((double)p[j].x + (double)p[i].x) * ((double)p[j].y - (double)p[i].y);
    j = i;
  }
  return -a * 0.5;
}

// src (n, cleaned + oriented in place) -> out raw offset path; returns count
__host__ __device__ int clipper_offset_raw(ipt *src, int n_in, double delta, ipt *out) {
  int hi = n_in - 1;
  while (hi > 0 && src[0].x == src[hi].x && src[0].y == src[hi].y) hi--;
  int n = 0;
  for (int i = 0; i <= hi; ++i)
    if (n == 0 || src[n - 1].x != src[i].x || src[n - 1].y != src[i].y) src[n++] = src[i];
  if (n < 3) return 0;
  if (!(clipper_area(src, n) >= 0))
    for (int i = 0, j = n - 1; i < j; ++i, --j) { ipt t = src[i]; src[i] = src[j]; src[j] = t; }
  int m = 0;
  if (fabs(delta) < 1.0e-20) {
    for (int i = 0; i < n; ++i) out[m++] = src[i];
    return m;
  }
  const double miter_lim = 0.5;
  // unit normal of edge j -> j+1
  auto normal = [&](int j, double &nx, double &ny) {
    ipt p1 = src[j], p2 = src[j + 1 == n ? 0 : j + 1];
    if (p1.x == p2.x && p1.y == p2.y) { nx = 0; ny = 0; return; }
    double dx = (double)(p2.x - p1.x), dy = (double)(p2.y - p1.y);
    double f = 1.0 / sqrt(dx * dx + dy * dy);
    dx *= f; dy *= f;
    nx = dy; ny = -dx;
  };
  double nkx, nky;
  normal(n - 1, nkx, nky);
  for (int j = 0; j < n; ++j) {
    double njx, njy;
    normal(j, njx, njy);
    double sinA = nkx * njy - njx * nky;
    bool done = false;
    if (fabs(sinA * delta) < 1.0) {
      double cosA = nkx * njx + njy * nky;
      if (cosA > 0) {
        out[m].x = (int)clip_round(src[j].x + nkx * delta);
        out[m].y = (int)clip_round(src[j].y + nky * delta);
        m++; done = true;
      }
    } else if (sinA > 1.0) sinA = 1.0;
    else if (sinA < -1.0) sinA = -1.0;
    if (!done) {
      if (sinA * delta < 0) {
        out[m].x = (int)clip_round(src[j].x + nkx * delta);
        out[m].y = (int)clip_round(src[j].y + nky * delta); m++;
        out[m++] = src[j];
        out[m].x = (int)clip_round(src[j].x + njx * delta);
        out[m].y = (int)clip_round(src[j].y + njy * delta); m++;
      } else {
        double r = 1 + (njx * nkx + njy * nky);
        if (r >= miter_lim) {
          double q = delta / r;
          out[m].x = (int)clip_round(src[j].x + (nkx + njx) * q);
          out[m].y = (int)clip_round(src[j].y + (nky + njy) * q); m++;
        } else {
          double dx = tan(atan2(sinA, nkx * njx + nky * njy) / 4);
          out[m].x = (int)clip_round(src[j].x + delta * (nkx - nky * dx));
          out[m].y = (int)clip_round(src[j].y + delta * (nky + nkx * dx)); m++;
          out[m].x = (int)clip_round(src[j].x + delta * (njx + njy * dx));
          out[m].y = (int)clip_round(src[j].y + delta * (njy - njx * dx)); m++;
        }
      }
    }
    nkx = njx; nky = njy;
  }
  return m;
}

__host__ __device__ void clipper_intersect_point(ipt a0, ipt a1, ipt b0, ipt b1, ipt *ip) {
  ipt abot, atop, bbot, btop;
  if (a0.y >= a1.y) { abot = a0; atop = a1; } else { abot = a1; atop = a0; }
  if (b0.y >= b1.y) { bbot = b0; btop = b1; } else { bbot = b1; btop = b0; }
  const double HORIZ = -1.0E+40;
  double adx = (atop.y == abot.y) ? HORIZ : (double)(atop.x - abot.x) / (double)(atop.y - abot.y);
  double bdx = (btop.y == bbot.y) ? HORIZ : (double)(btop.x - bbot.x) / (double)(btop.y - bbot.y);
  double b1_, b2_;
  long long X, Y;
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
  ip->x = (int)X; ip->y = (int)Y;
}

struct rat { long long num, den; };
__host__ __device__ __forceinline__ bool rat_lt(rat a, rat b) { return (__int128)a.num * b.den < (__int128)b.num * a.den; }
__host__ __device__ __forceinline__ bool rat_eq(rat a, rat b) { return (__int128)a.num * b.den == (__int128)b.num * a.den; }

__host__ __device__ bool seg_hit(const ipt *Q, int m, int i, int j, int which, rat *t, rat *s) {
  ipt a0 = Q[i], a1 = Q[i + 1 == m ? 0 : i + 1], b0 = Q[j], b1 = Q[j + 1 == m ? 0 : j + 1];
  long long dix = a1.x - a0.x, diy = a1.y - a0.y, djx = b1.x - b0.x, djy = b1.y - b0.y;
  long long wx = b0.x - a0.x, wy = b0.y - a0.y;
  long long den = crossi(dix, diy, djx, djy);
  if (den != 0) {
    if (which != 0) return false;
    long long tn = crossi(wx, wy, djx, djy), sn = crossi(wx, wy, dix, diy);
    if (den < 0) { den = -den; tn = -tn; sn = -sn; }
    if (tn < 0 || tn > den || sn < 0 || sn > den) return false;
    t->num = tn; t->den = den; s->num = sn; s->den = den;
    return true;
  }
  if (crossi(wx, wy, dix, diy) != 0) return false;
  long long L = doti(dix, diy, dix, diy);
  if (L == 0 || which == 0) return false;
  ipt e = which == 1 ? b0 : b1;
  long long tn = doti(e.x - a0.x, e.y - a0.y, dix, diy);
  if (tn < 0 || tn > L) return false;
  t->num = tn; t->den = L; s->num = which == 1 ? 0 : 1; s->den = 1;
  return true;
}

__host__ __device__ __forceinline__ int half_of(long long rx, long long ry, long long dx, long long dy) {
  long long c = crossi(rx, ry, dx, dy), d = doti(rx, ry, dx, dy);
  if (c > 0) return 0;
  if (c < 0) return 2;
  return d < 0 ? 1 : 3;
}
__host__ __device__ __forceinline__ bool ccw_before(long long rx, long long ry, long long ax, long long ay, long long bx, long long by) {
  int ha = half_of(rx, ry, ax, ay), hb = half_of(rx, ry, bx, by);
  if (ha != hb) return ha < hb;
  if (ha == 1 || ha == 3) return false;
  return crossi(ax, ay, bx, by) > 0;
}
__host__ __device__ __forceinline__ bool same_dir(long long ax, long long ay, long long bx, long long by) {
  return crossi(ax, ay, bx, by) == 0 && doti(ax, ay, bx, by) > 0;
}

// Clipper's clean-up of one closed offset path = the boundary of {winding > 0}, walked with the region on
// the left; the derivation, the node rule and the "first surviving polygon" rule are spelled out at
// orc_union_positive in oracle/postproc_oracle.c, which this code must match vertex for vertex.  The
// restatement is pinned by the reference's golden polygons and ground-truth maps (tests/test_oracle_goldens.py)
// and, independently of both implementations, by a winding-number rasteriser (oracle/region_check.c).
struct ray_t { int dx, dy, sign, seg; rat s; };  // part of path segment `seg` leaving (+1) / reaching (-1) a node
constexpr int MAX_RAYS = 32;                    // more segments through one point: the candidate is dropped

// all rays at P = (pxn, pyn) / pden; returns the count, -1 on overflow
__host__ __device__ int rays_at(const ipt *Q, int m, __int128 pxn, __int128 pyn, long long pden, ray_t *rays, bool *is_vertex, ipt *vtx) {
  int k = 0;
  *is_vertex = false;
  // P in floating point with a one-pixel margin rejects almost every segment before the exact test
  const double fx = (double)pxn / (double)pden, fy = (double)pyn / (double)pden;
  for (int j = 0; j < m; ++j) {
    const ipt b0 = Q[j], b1 = Q[j + 1 == m ? 0 : j + 1];
    if ((b0.x < fx - 1 && b1.x < fx - 1) || (b0.x > fx + 1 && b1.x > fx + 1) || (b0.y < fy - 1 && b1.y < fy - 1) || (b0.y > fy + 1 && b1.y > fy + 1)) continue;
    const long long dx = (long long)b1.x - b0.x, dy = (long long)b1.y - b0.y;
    const __int128 qx = pxn - (__int128)b0.x * pden, qy = pyn - (__int128)b0.y * pden;  // (P - b0) * pden
    if (qx * dy - qy * dx != 0) continue;
    // parameter of P along the segment from its dominant coordinate (64-bit safe for 16-bit coordinates)
    const bool use_x = (dx < 0 ? -dx : dx) >= (dy < 0 ? -dy : dy);
    const long long dd = use_x ? dx : dy;
    __int128 sn = use_x ? qx : qy;
    if (dd < 0) sn = -sn;
    const __int128 sd = (__int128)pden * (dd < 0 ? -dd : dd);
    if (sn < 0 || sn > sd) continue;
    if (sn == 0) { *is_vertex = true; *vtx = b0; }
    if (sn == sd) { *is_vertex = true; *vtx = b1; }
    const rat s = {(long long)sn, (long long)sd};
    if (sn < sd) { if (k >= MAX_RAYS) return -1; rays[k].dx = (int)dx; rays[k].dy = (int)dy; rays[k].sign = 1; rays[k].seg = j; rays[k].s = s; k++; }
    if (sn > 0) { if (k >= MAX_RAYS) return -1; rays[k].dx = (int)-dx; rays[k].dy = (int)-dy; rays[k].sign = -1; rays[k].seg = j; rays[k].s = s; k++; }
  }
  return k;
}

// turn counter-clockwise about the node from just after direction r (winding w0 there); the first group of
// coincident rays across which the winding becomes positive carries the boundary on.  Returns an outgoing ray
// of that group (-1: none) and the winding on its right.
__host__ __device__ int next_boundary_ray(ray_t *rays, int k, long long rx, long long ry, int w0, int prefer_seg, int *w_right) {
  for (int i = 1; i < k; ++i) {
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

// winding number at (P.x - eps, P.y + delta), 0 < eps << delta << 1
__host__ __device__ int winding_above_left(const ipt *Q, int m, __int128 pxn, __int128 pyn, long long pden) {
  int w = 0;
  for (int j = 0; j < m; ++j) {
    const ipt a = Q[j], b = Q[j + 1 == m ? 0 : j + 1];
    const __int128 ax = (__int128)a.x * pden, bx = (__int128)b.x * pden;
    int dir;
    if (ax < pxn && pxn <= bx) dir = -1;
    else if (bx < pxn && pxn <= ax) dir = 1;
    else continue;
    const __int128 lhs = (pxn - ax) * ((long long)b.y - a.y), rhs = (pyn - (__int128)a.y * pden) * ((long long)b.x - a.x);
    const bool above = (b.x > a.x) ? (lhs > rhs) : (lhs < rhs);
    if (above) w += dir;
  }
  return w;
}

struct rpt { __int128 xn, yn; long long den; };
__host__ __device__ __forceinline__ bool node_after(const rpt &a, const rpt &b) {  // a strictly after b in (y descending, x ascending)
  const __int128 ya = a.yn * b.den, yb = b.yn * a.den;
  if (ya != yb) return ya < yb;
  return a.xn * b.den > b.xn * a.den;
}

// one ring from (start_seg, start_t), then FixupOutPolygon; < 3: collapsed, -1: failure
__host__ __device__ int walk_ring(const ipt *Q, int m, int start_seg, rat start_t, int w_right, ipt *out, int cap, ray_t *rays) {
  int cur = start_seg, n_out = 0;
  rat cur_t = start_t;
  int guard = 0;
  const int max_iter = 8 * m + 64;
  for (;;) {
    if (++guard > max_iter) return -1;
    rat t_best = {1, 1};
    // bounding box of the current segment: a segment whose (closed) box misses it cannot hit it
    // in any of the three ways seg_hit distinguishes — an exact, cheap reject of most pairs
    const ipt c0 = Q[cur], c1 = Q[cur + 1 == m ? 0 : cur + 1];
    const int cminx = c0.x < c1.x ? c0.x : c1.x, cmaxx = c0.x < c1.x ? c1.x : c0.x;
    const int cminy = c0.y < c1.y ? c0.y : c1.y, cmaxy = c0.y < c1.y ? c1.y : c0.y;
    for (int j = 0; j < m; ++j) {
      if (j == cur) continue;
      const ipt b0 = Q[j], b1 = Q[j + 1 == m ? 0 : j + 1];
      if ((b0.x < cminx && b1.x < cminx) || (b0.x > cmaxx && b1.x > cmaxx) || (b0.y < cminy && b1.y < cminy) || (b0.y > cmaxy && b1.y > cmaxy)) continue;
      for (int which = 0; which < 3; ++which) {
        rat t, s;
        if (!seg_hit(Q, m, cur, j, which, &t, &s)) continue;
        if (!rat_lt(cur_t, t)) continue;
        if (rat_lt(t, t_best)) t_best = t;
      }
    }
    const long long ux = (long long)c1.x - c0.x, uy = (long long)c1.y - c0.y;
    const long long pden = t_best.den;
    const __int128 pxn = (__int128)c0.x * pden + (__int128)t_best.num * ux, pyn = (__int128)c0.y * pden + (__int128)t_best.num * uy;
    bool node_is_vertex;
    ipt node_v = {0, 0};
    const int k = rays_at(Q, m, pxn, pyn, pden, rays, &node_is_vertex, &node_v);
    if (k < 0) return -1;
    int wr;
    const int pick = next_boundary_ray(rays, k, -ux, -uy, w_right, cur, &wr);
    if (pick < 0) return -1;
    const int nxt = rays[pick].seg;
    const rat nxt_s = rays[pick].s;
    if (node_is_vertex || nxt != cur) {
      ipt node;
      if (node_is_vertex) node = node_v;
      else clipper_intersect_point(c0, c1, Q[nxt], Q[nxt + 1 == m ? 0 : nxt + 1], &node);
      if (n_out >= cap) return -1;
      out[n_out++] = node;
    }
    if (nxt == start_seg && rat_eq(nxt_s, start_t)) break;  // closed the ring (the start node was emitted last)
    cur = nxt; cur_t = nxt_s; w_right = wr;
  }
  // FixupOutPolygon: drop duplicates and collinear middles until stable
  bool changed = true;
  while (changed && n_out >= 3) {
    changed = false;
    for (int i = 0; i < n_out && n_out >= 3; ++i) {
      ipt p = out[(i + n_out - 1) % n_out], c = out[i], nn = out[(i + 1) % n_out];
      bool dup = (c.x == nn.x && c.y == nn.y) || (c.x == p.x && c.y == p.y);
      bool col = crossi((long long)c.x - p.x, (long long)c.y - p.y, (long long)nn.x - c.x, (long long)nn.y - c.y) == 0;
      if (dup || col) {
        for (int k2 = i; k2 + 1 < n_out; ++k2) out[k2] = out[k2 + 1];
        n_out--; changed = true; i--;
      }
    }
  }
  return n_out;
}

// Q: scratch for the de-duplicated path (also the rotation scratch).
__host__ __device__ int union_positive(const ipt *Qin, int m_in, ipt *Q, ipt *out, int cap, ipt *fast, int fast_cap) {
  int m = 0;
  for (int i = 0; i < m_in; ++i)
    if (m == 0 || Q[m - 1].x != Qin[i].x || Q[m - 1].y != Qin[i].y) Q[m++] = Qin[i];
  while (m > 1 && Q[0].x == Q[m - 1].x && Q[0].y == Q[m - 1].y) m--;
  if (m < 3) return 0;
  // the walk below reads the vertex list thousands of times: keep it in the caller's fast
  // (shared-memory) scratch when it fits
  ipt *Qg = Q;
  if (fast && m <= fast_cap) {
    for (int i = 0; i < m; ++i) fast[i] = Q[i];
    Q = fast;
  }
  ray_t rays[MAX_RAYS];
  // start search in (y descending, x ascending) order: the largest-y vertex first (nearly always on the
  // boundary), then all vertices and crossings
  rpt last = {0, 0, 1};
  int n_out = 0;
  for (int tries = 0; tries < 4 * m + 16; ++tries) {
    rpt best = {0, 0, 0};
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
          const int i1 = i + 1 == m ? 0 : i + 1;
          const long long ux = (long long)Q[i1].x - Q[i].x, uy = (long long)Q[i1].y - Q[i].y;
          rpt c = {(__int128)Q[i].x * t.den + (__int128)t.num * ux, (__int128)Q[i].y * t.den + (__int128)t.num * uy, t.den};
          if (!node_after(c, last)) continue;
          if (best.den == 0 || node_after(best, c)) best = c;
        }
    }
    if (best.den == 0) break;
    last = best;
    bool isv;
    ipt vt;
    const int k = rays_at(Q, m, best.xn, best.yn, best.den, rays, &isv, &vt);
    if (k < 0) break;
    if (k == 0) continue;
    const int w0 = winding_above_left(Q, m, best.xn, best.yn, best.den);
    int w_right;
    const int pick = next_boundary_ray(rays, k, 0, 1, w0, -1, &w_right);
    if (pick < 0) continue;
    n_out = walk_ring(Q, m, rays[pick].seg, rays[pick].s, w_right, out, cap, rays);
    if (n_out < 0) { n_out = 0; break; }
    if (n_out >= 3) break;
    n_out = 0;
  }
  if (n_out < 3) return 0;
  // BuildResult order: start right after the last top-most vertex (rotate in place via Q)
  int top = 0;
  for (int i = 1; i < n_out; ++i)
    if (out[i].y < out[top].y || (out[i].y == out[top].y && out[i].x > out[top].x)) top = i;
  int st = (top + 1) % n_out;
  for (int i = 0; i < n_out; ++i) Qg[i] = out[(st + i) % n_out];  // n_out <= cap = capacity of the Q slab region
  for (int i = 0; i < n_out; ++i) out[i] = Qg[i];
  return n_out;
}

// ---- min-area rectangle (imageproc 0.22 min_area_rect + metrics.rs:133-148) -------------
__host__ __device__ __forceinline__ int orient(dpt p, dpt q, dpt r) {
  double val = (q.y - p.y) * (r.x - q.x) - (q.x - p.x) * (r.y - q.y);
  if (val == 0.0) return 0;
  return val > 0.0 ? 1 : 2;
}
__host__ __device__ __forceinline__ double ddist(dpt a, dpt b) { return sqrt((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y)); }
__host__ __device__ __forceinline__ dpt rot(dpt p, double s, double c) { dpt r; r.x = p.x * c + p.y * s; r.y = p.y * c - p.x * s; return r; }
__host__ __device__ __forceinline__ dpt irot(dpt p, double s, double c) { dpt r; r.x = p.x * c - p.y * s; r.y = p.y * c + p.x * s; return r; }
__host__ __device__ __forceinline__ double pt_dist(ipt a, ipt b) {
  double dx = (double)a.x - (double)b.x, dy = (double)a.y - (double)b.y;
  return sqrt(dx * dx + dy * dy);
}

// work: >= n points, hull: >= n+1 points
__host__ __device__ double min_area_bounding_box(const ipt *pts, int n, dpt *work, dpt *hull, ipt box_out[4]) {
  ipt b[4];
  for (int i = 0; i < n; ++i) { work[i].x = pts[i].x; work[i].y = pts[i].y; }
  int s = 0;
  for (int i = 1; i < n; ++i)
    if (work[i].y < work[s].y || (work[i].y == work[s].y && work[i].x < work[s].x)) s = i;
  dpt start = work[s];
  work[s] = work[0];
  dpt *rest = work + 1;
  int nr = n - 1;
  for (int i = 1; i < nr; ++i) {  // stable insertion sort by polar order around `start`
    dpt key = rest[i];
    int j = i - 1;
    while (j >= 0) {
      int o = orient(start, key, rest[j]);
      bool less = (o == 0) ? (ddist(start, key) < ddist(start, rest[j])) : (o == 2);
      if (!less) break;
      rest[j + 1] = rest[j];
      j--;
    }
    rest[j + 1] = key;
  }
  int nrem = 0;
  for (int i = 0; i < nr;) {
    int k = i;
    while (k + 1 < nr && orient(start, rest[k], rest[k + 1]) == 0) k++;
    rest[nrem++] = rest[k];
    i = k + 1;
  }
  int h = 0;
  hull[h++] = start;
  for (int i = 0; i < nrem; ++i) {
    while (h > 1 && orient(hull[h - 2], hull[h - 1], rest[i]) != 2) h--;
    hull[h++] = rest[i];
  }
  if (h == 1) {
    for (int i = 0; i < 4; ++i) { b[i].x = (int)hull[0].x; b[i].y = (int)hull[0].y; }
  } else if (h == 2) {
    b[0].x = (int)hull[0].x; b[0].y = (int)hull[0].y;
    b[1].x = (int)hull[1].x; b[1].y = (int)hull[1].y;
    b[2] = b[1]; b[3] = b[0];
  } else {
    const double PI = 3.14159265358979323846264338327950288;
    double min_area = 1.7976931348623157e308;
    dpt res[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    for (int e = 0; e + 1 < h; ++e) {
      double ex = hull[e + 1].x - hull[e].x, ey = hull[e + 1].y - hull[e].y;
      // correctly rounded atan2 / sin / cos (dd_math.cuh): the reference's libm (glibc) rounds
      // correctly in ~99.9 % of calls, CUDA's libm is 1-2 ulp off far more often, and one ulp
      // flips the outward floor/ceil below whenever a rotated coordinate is an exact integer
      double angle = fabs(fmod(ddm::cr_atan2(ey, ex) + PI, PI / 2.));
      double sn, cs;
      ddm::cr_sincos(angle, &sn, &cs);
      double min_x = 1.7976931348623157e308, max_x = -1.7976931348623157e308;
      double min_y = 1.7976931348623157e308, max_y = -1.7976931348623157e308;
      for (int i = 0; i < h; ++i) {
        dpt r = rot(hull[i], sn, cs);
        if (r.x < min_x) min_x = r.x;
        if (r.x > max_x) max_x = r.x;
        if (r.y < min_y) min_y = r.y;
        if (r.y > max_y) max_y = r.y;
      }
      double area = (max_x - min_x) * (max_y - min_y);
      if (area < min_area) {
        min_area = area;
        dpt a = {max_x, min_y}, bb = {min_x, min_y}, cc = {min_x, max_y}, d = {max_x, max_y};
        res[0] = irot(a, sn, cs); res[1] = irot(bb, sn, cs); res[2] = irot(cc, sn, cs); res[3] = irot(d, sn, cs);
      }
    }
    for (int i = 1; i < 4; ++i) {
      dpt key = res[i]; int j = i - 1;
      while (j >= 0 && key.x < res[j].x) { res[j + 1] = res[j]; j--; }
      res[j + 1] = key;
    }
    int i1 = res[1].y > res[0].y ? 0 : 1;
    int i2 = res[3].y > res[2].y ? 2 : 3;
    int i3 = res[3].y > res[2].y ? 3 : 2;
    int i4 = res[1].y > res[0].y ? 1 : 0;
    b[0].x = (int)floor(res[i1].x); b[0].y = (int)floor(res[i1].y);
    b[1].x = (int)ceil(res[i2].x);  b[1].y = (int)floor(res[i2].y);
    b[2].x = (int)ceil(res[i3].x);  b[2].y = (int)ceil(res[i3].y);
    b[3].x = (int)floor(res[i4].x); b[3].y = (int)ceil(res[i4].y);
  }
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
  if (box_out) for (int i = 0; i < 4; ++i) box_out[i] = r[i];
  double w = pt_dist(r[0], r[1]), hh = pt_dist(r[0], r[3]);
  return w < hh ? w : hh;
}

// slab layout per candidate (units of 8 bytes), n = DP vertex count:
//   src  [n+1]         int2   DP polygon copy (cleaned / oriented in place)
//   raw  [3n+3]        int2   raw offset path
//   Q    [max(3n+3, cap)] int2 de-duplicated path / rotation scratch
//   out  [cap]         int2   expanded polygon, cap = 6n+32
//   work [cap] hull [cap+1]   double2 (2 units each)
__host__ __device__ inline int unclip_cap(int n) { return 6 * n + 32; }
__host__ __device__ inline int64_t unclip_slab_units(int n) {
  int64_t cap = unclip_cap(n);
  return (n + 1) + (3 * n + 3) + cap + cap + 2 * cap + 2 * (cap + 1);
}

__global__ void unclip_slab_size_kernel(const int *__restrict__ cand_contour, const int *__restrict__ dp_count, int n_cand,
                                        int64_t *__restrict__ units) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cand) return;
  int n = dp_count[cand_contour ? cand_contour[i] : i];
  units[i] = unclip_slab_units(n);
}

// status: 0 dropped by score, 1 kept, 2 dropped (empty offset, reference panics, D11),
//         3 dropped by min_size
constexpr int UNCLIP_THREADS = 64;
constexpr int UNCLIP_FAST_PTS = 256;  // 2 KB of shared memory per active lane

__global__ void __launch_bounds__(UNCLIP_THREADS) unclip_kernel(const int *__restrict__ cand_contour, const int64_t *__restrict__ chain_off,
                              const ushort2 *__restrict__ dp_pts, const int *__restrict__ dp_count, int n_cand,
                              const double *__restrict__ scores, double box_thresh, double min_size, double factor,
                              const int64_t *__restrict__ slab_off, int2 *__restrict__ slabs, int *__restrict__ out_count,
                              uint8_t *__restrict__ status, double *__restrict__ sside_out, int2 *__restrict__ box_out) {
  const int64_t si = sparse_item_index();
  if (si < 0 || si >= n_cand) return;
  const int i = (int)si;
  out_count[i] = 0;
  double score = scores[i];
  if (score < 0.0 || box_thresh > score) {  // metrics.rs:100 (`score < 0` marks a rejected candidate)
    status[i] = 0;
    return;
  }
  const int c = cand_contour ? cand_contour[i] : i;
  const int n = dp_count[c];
  const ushort2 *pts = dp_pts + chain_off[c];
  const int cap = unclip_cap(n);
  ipt *src = reinterpret_cast<ipt *>(slabs + slab_off[i]);
  ipt *raw = src + (n + 1);
  ipt *Q = raw + (3 * n + 3);
  ipt *out = Q + cap;
  dpt *work = reinterpret_cast<dpt *>(out + cap);
  dpt *hull = work + cap;
  // geo: unsigned_area and euclidean_length of the closed ring
  double twice = 0.0, perim = 0.0;
  for (int k = 0; k < n; ++k) {
    ushort2 a = pts[k], b = pts[k + 1 == n ? 0 : k + 1];
    twice += (double)a.x * (double)b.y - (double)a.y * (double)b.x;
    ipt ia = {a.x, a.y}, ib = {b.x, b.y};
    perim += pt_dist(ia, ib);
    src[k] = ia;
  }
  double area = fabs(twice / 2.0);
  double distance = area * factor / perim;
  int m = clipper_offset_raw(src, n, distance, raw);
  __shared__ ipt s_fast[(UNCLIP_THREADS / 32) * SPARSE_LANES][UNCLIP_FAST_PTS];
  int ne = m >= 3 ? union_positive(raw, m, Q, out, cap, s_fast[(threadIdx.x >> 5) * SPARSE_LANES + (threadIdx.x & 31)], UNCLIP_FAST_PTS) : 0;
  if (ne == 0) { status[i] = 2; return; }
  ipt box[4];
  double sside = min_area_bounding_box(out, ne, work, hull, box);
  if (sside_out) sside_out[i] = sside;
  if (box_out) for (int k = 0; k < 4; ++k) box_out[i * 4 + k] = make_int2(box[k].x, box[k].y);
  if (sside < min_size) { status[i] = 3; out_count[i] = ne; return; }
  status[i] = 1;
  out_count[i] = ne;
}

// polygon::clip_polygon (polygon.rs:13-49) on the HOST, from the very functions the unclip kernel runs (they are
// __host__ __device__): geo's unsigned_area / euclidean_length -> signed distance -> Clipper offset (miter 2) -> union
// clean-up -> first polygon.  shrink_polygon (polygon.rs:44-49) is host code in the reference too — it prepares the
// training targets (image_ops.rs:222-277) — and is not part of the GPU path; expand_polygon's product form stays
// ocrb_expand_polygon (device).  Besides completing the polygon.rs mirror, this lets the CPU test-suite hold the shipped
// offset / union source to the oracle and to the reference's gt_shrinked fixtures without a GPU.
int clip_polygon_host(const int32_t *xy, int n, double factor, int shrink, int32_t *out_xy, int cap_pts, int *n_out, double *distance_out) {
  *n_out = 0;
  if (n < 1) return OCRB_OK;
  const int cap = unclip_cap(n);
  std::vector<ipt> src((size_t)n + 1), raw((size_t)3 * n + 3), Q((size_t)cap), out((size_t)cap);
  double twice = 0.0, perim = 0.0;
  for (int k = 0; k < n; ++k) {
    const ipt a = {xy[2 * k], xy[2 * k + 1]};
    const int k1 = k + 1 == n ? 0 : k + 1;
    const ipt b = {xy[2 * k1], xy[2 * k1 + 1]};
    twice += (double)a.x * (double)b.y - (double)a.y * (double)b.x;
    perim += pt_dist(a, b);
    src[k] = a;
  }
  const double area = fabs(twice / 2.0);
  double distance = area * factor / perim;
  if (shrink) distance *= -1.;
  if (distance_out) *distance_out = distance;
  const int m = clipper_offset_raw(src.data(), n, distance, raw.data());
  const int ne = m >= 3 ? union_positive(raw.data(), m, Q.data(), out.data(), cap, nullptr, 0) : 0;
  if (ne > cap_pts) { set_error("clip_polygon: %d points, room for %d", ne, cap_pts); return OCRB_ERR_CAPACITY; }
  for (int k = 0; k < ne; ++k) { out_xy[2 * k] = out[k].x; out_xy[2 * k + 1] = out[k].y; }
  *n_out = ne;
  return OCRB_OK;
}

// host-only test hook (ocrb_debug_min_area_bounding_box_host): get_min_area_bounding_box (metrics.rs:133-148) computed on the
// HOST by the very function the unclip kernel runs — the CPU suite holds the shipped source to the reference's known answer
// and to the oracle.  The product entry point is ocrb_min_area_bounding_box (device).
int min_area_bounding_box_host(const int32_t *xy, int n, int32_t *box_xy, double *sside) {
  std::vector<ipt> pts((size_t)n);
  for (int k = 0; k < n; ++k) pts[k] = {xy[2 * k], xy[2 * k + 1]};
  std::vector<dpt> work((size_t)n + 1), hull((size_t)n + 2);
  ipt box[4];
  const double s = min_area_bounding_box(pts.data(), n, work.data(), hull.data(), box);
  for (int k = 0; k < 4; ++k) { box_xy[2 * k] = box[k].x; box_xy[2 * k + 1] = box[k].y; }
  *sside = s;
  return OCRB_OK;
}

int launch_unclip_slab_sizes(ocrb_ctx *ctx, const int *cand_contour, const int *dp_count, int n_cand, int64_t *units) {
  if (n_cand <= 0) return OCRB_OK;
  unclip_slab_size_kernel<<<(unsigned)cdiv(n_cand, 128), 128, 0, ctx->stream>>>(cand_contour, dp_count, n_cand, units);
  return check_launch(ctx, "unclip_slab_size");
}

int launch_unclip(ocrb_ctx *ctx, const int *cand_contour, const int64_t *chain_off, const ushort2 *dp_pts,
                  const int *dp_count, int n_cand, const double *scores, double box_thresh, double min_size, double factor,
                  const int64_t *slab_off, int2 *slabs, int *out_count, uint8_t *status, double *sside_out, int2 *box_out) {
  if (n_cand <= 0) return OCRB_OK;
  unclip_kernel<<<sparse_grid(n_cand, UNCLIP_THREADS), UNCLIP_THREADS, 0, ctx->stream>>>(cand_contour, chain_off, dp_pts, dp_count, n_cand, scores,
                                                                    box_thresh, min_size, factor, slab_off, slabs, out_count,
                                                                    status, sside_out, box_out);
  return check_launch(ctx, "unclip");
}

// expanded polygon location inside a candidate's slab
__device__ __forceinline__ const int2 *slab_out_ptr(const int2 *slabs, int64_t off, int n) {
  return slabs + off + (n + 1) + (3 * n + 3) + unclip_cap(n);
}

// kept polygons -> result arrays (metrics.rs:109-123: coordinate / adjust, round, as u32)
__device__ __forceinline__ uint32_t sat_u32(double v) {
  if (!(v == v)) return 0;
  if (v <= 0.0) return 0;
  if (v >= 4294967295.0) return 4294967295u;
  return (uint32_t)v;
}

__global__ void emit_polygons_kernel(const int *__restrict__ cand_contour, const int *__restrict__ dp_count,
                                     const int64_t *__restrict__ start_idx, int64_t image_stride, int n_cand,
                                     const uint8_t *__restrict__ status, const int *__restrict__ kept_rank,
                                     const int64_t *__restrict__ pt_off, const int64_t *__restrict__ slab_off,
                                     const int2 *__restrict__ slabs, const int *__restrict__ out_count,
                                     const double *__restrict__ scores, const double *__restrict__ adjust,
                                     uint32_t *__restrict__ xy, double *__restrict__ out_scores,
                                     int64_t *__restrict__ out_pt_off, int *__restrict__ out_image,
                                     const int2 *__restrict__ cand_box, int2 *__restrict__ out_box) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cand || status[i] != 1) return;
  const int c = cand_contour ? cand_contour[i] : i;
  const int n = dp_count[c];
  const int64_t b = start_idx ? start_idx[c] / image_stride : 0;
  const double ax = adjust[b * 2], ay = adjust[b * 2 + 1];
  const int2 *src = slab_out_ptr(slabs, slab_off[i], n);
  const int r = kept_rank[i];
  const int64_t po = pt_off[i];
  const int ne = out_count[i];
  for (int k = 0; k < ne; ++k) {
    xy[2 * (po + k)] = sat_u32(round((double)src[k].x / ax));
    xy[2 * (po + k) + 1] = sat_u32(round((double)src[k].y / ay));
  }
  out_scores[r] = scores[i];
  out_pt_off[r] = po;
  out_image[r] = (int)b;
  if (out_box)
    for (int k = 0; k < 4; ++k) out_box[r * 4 + k] = cand_box[i * 4 + k];
}

int launch_emit_polygons(ocrb_ctx *ctx, const int *cand_contour, const int *dp_count, const int64_t *start_idx,
                         int64_t image_stride, int n_cand, const uint8_t *status, const int *kept_rank,
                         const int64_t *pt_off, const int64_t *slab_off, const int2 *slabs, const int *out_count,
                         const double *scores, const double *adjust, uint32_t *xy, double *out_scores,
                         int64_t *out_pt_off, int *out_image, const int2 *cand_box, int2 *out_box) {
  if (n_cand <= 0) return OCRB_OK;
  emit_polygons_kernel<<<(unsigned)cdiv(n_cand, 128), 128, 0, ctx->stream>>>(cand_contour, dp_count, start_idx, image_stride,
                                                                            n_cand, status, kept_rank, pt_off, slab_off, slabs,
                                                                            out_count, scores, adjust, xy, out_scores,
                                                                            out_pt_off, out_image, cand_box, out_box);
  return check_launch(ctx, "emit_polygons");
}

// helper kernels for the compaction steps
__global__ void flag_ge4_kernel(const int *__restrict__ dp_count, int64_t n, uint8_t *__restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = dp_count[i] >= 4 ? 1 : 0;
}
__global__ void compact_index_kernel(const uint8_t *__restrict__ flag, const int *__restrict__ rank, int64_t n, int *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) out[rank[i]] = (int)i;
}
__global__ void kept_sizes_kernel(const uint8_t *__restrict__ status, const int *__restrict__ out_count, int n,
                                  uint8_t *__restrict__ kept_flag, int *__restrict__ kept_pts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool k = status[i] == 1;
  kept_flag[i] = k ? 1 : 0;
  kept_pts[i] = k ? out_count[i] : 0;
}
// per-image statistics: contours, >=4 dp points, >= box_thresh, kept, dropped(empty offset)
__global__ void stats_contours_kernel(const int64_t *__restrict__ start_idx, const int *__restrict__ dp_count, int64_t n,
                                      int64_t image_stride, unsigned long long *__restrict__ stats) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t b = start_idx[i] / image_stride;
  atomicAdd(&stats[b * 5 + 0], 1ull);
  if (dp_count[i] >= 4) atomicAdd(&stats[b * 5 + 1], 1ull);
}
__global__ void stats_cands_kernel(const int *__restrict__ cand_contour, const int64_t *__restrict__ start_idx, int n,
                                   int64_t image_stride, const uint8_t *__restrict__ status,
                                   unsigned long long *__restrict__ stats) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t b = start_idx[cand_contour[i]] / image_stride;
  uint8_t s = status[i];
  if (s != 0) atomicAdd(&stats[b * 5 + 2], 1ull);
  if (s == 1) atomicAdd(&stats[b * 5 + 3], 1ull);
  if (s == 2) atomicAdd(&stats[b * 5 + 4], 1ull);
}

int launch_flag_ge4(ocrb_ctx *ctx, const int *dp_count, int64_t n, uint8_t *flag) {
  if (n <= 0) return OCRB_OK;
  flag_ge4_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(dp_count, n, flag);
  return check_launch(ctx, "flag_ge4");
}
int launch_compact_index(ocrb_ctx *ctx, const uint8_t *flag, const int *rank, int64_t n, int *out) {
  if (n <= 0) return OCRB_OK;
  compact_index_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(flag, rank, n, out);
  return check_launch(ctx, "compact_index");
}
int launch_kept_sizes(ocrb_ctx *ctx, const uint8_t *status, const int *out_count, int n, uint8_t *kept_flag, int *kept_pts) {
  if (n <= 0) return OCRB_OK;
  kept_sizes_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(status, out_count, n, kept_flag, kept_pts);
  return check_launch(ctx, "kept_sizes");
}
int launch_stats(ocrb_ctx *ctx, const int64_t *start_idx, const int *dp_count, int64_t n_contours, int64_t image_stride,
                 const int *cand_contour, int n_cand, const uint8_t *status, unsigned long long *stats) {
  if (n_contours > 0) {
    stats_contours_kernel<<<(unsigned)cdiv(n_contours, 256), 256, 0, ctx->stream>>>(start_idx, dp_count, n_contours, image_stride, stats);
    OCRB_TRY(check_launch(ctx, "stats_contours"));
  }
  if (n_cand > 0) {
    stats_cands_kernel<<<(unsigned)cdiv(n_cand, 256), 256, 0, ctx->stream>>>(cand_contour, start_idx, n_cand, image_stride, status, stats);
    OCRB_TRY(check_launch(ctx, "stats_cands"));
  }
  return OCRB_OK;
}

}  // namespace ocrb

// ---- single-polygon test hooks (ocrb_min_area_bounding_box) ----------------------------
namespace ocrb {
__global__ void minrect_hook_kernel(const int2 *__restrict__ pts, int n, double2 *work, double2 *hull, int2 *box, double *sside) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  ipt b[4];
  double s = min_area_bounding_box(reinterpret_cast<const ipt *>(pts), n, reinterpret_cast<dpt *>(work),
                                   reinterpret_cast<dpt *>(hull), b);
  for (int k = 0; k < 4; ++k) box[k] = make_int2(b[k].x, b[k].y);
  *sside = s;
}

int launch_minrect_hook(ocrb_ctx *ctx, const int2 *pts, int n, double2 *work, double2 *hull, int2 *box, double *sside) {
  minrect_hook_kernel<<<1, 32, 0, ctx->stream>>>(pts, n, work, hull, box, sside);
  return check_launch(ctx, "minrect_hook");
}
}  // namespace ocrb
