/*
 * ORACLE-SIDE CHECKER — test infrastructure only (tests/test_union_region.py).
 *
 * An independent statement of what Clipper's clean-up of an offset path must produce, sharing
 * no code and no method with orc_union_positive (oracle/postproc_oracle.c) or its CUDA
 * counterpart: the winding number of a closed integer path sampled on a regular sub-pixel
 * grid by plain scanline accumulation.  The test rasterises the RAW offset path (what
 * ClipperOffset::DoOffset emits, polygon.rs:27-31) and the polygon the implementation returns,
 * and requires {winding > 0} of the first to coincide with the inside of the second, up to the
 * slivers that rounding a crossing point to the integer grid creates.
 *
 * Sample (i, k) sits at x = x0 + (2 i + 1) / (2 S), y = y0 + (2 k + 1) / (2 S): never on a
 * vertex, never on a horizontal edge.  All comparisons are integer (scaled by 2 S), so a sample
 * exactly on an edge is classified the same way for every edge on that line.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int32_t x, y; } ipt;

/* out[k * nx + i] = winding number (counter-clockwise positive in the raw x,y plane) */
void rc_winding_raster(const ipt *p, int n, int x0, int y0, int S, int nx, int ny, int16_t *out) {
  const int64_t D = 2 * (int64_t)S;
  int32_t *diff = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nx + 1));
  for (int k = 0; k < ny; ++k) {
    memset(diff, 0, sizeof(int32_t) * (size_t)(nx + 1));
    const int64_t ys = (int64_t)y0 * D + 2 * k + 1; /* sample row, scaled by D */
    for (int e = 0; e < n; ++e) {
      ipt a = p[e], b = p[(e + 1) % n];
      if (a.y == b.y) continue;
      int sign = 1; /* upward (y increasing) edge: points on its LEFT (smaller x ... see below) */
      if (a.y > b.y) { ipt t = a; a = b; b = t; sign = -1; }
      const int64_t ay = (int64_t)a.y * D, by = (int64_t)b.y * D;
      if (!(ay < ys && ys < by)) continue; /* ys is odd, ay/by even: never equal */
      /* crossing x (scaled by D): xc = a.x*D + (ys - ay) * dx / dy; a sample at xs lies on the
       * +x side of the edge iff xs * dy >= a.x*D*dy + (ys - ay) * dx  (dy > 0 after the swap) */
      const int64_t dx = (int64_t)b.x - a.x, dy = (int64_t)b.y - a.y;
      const int64_t rhs = (int64_t)a.x * D * dy + (ys - ay) * dx;
      /* smallest i with (x0*D + 2i + 1) * dy >= rhs */
      int64_t num = rhs - ((int64_t)x0 * D + 1) * dy; /* 2 i dy >= num */
      int64_t i0;
      if (num <= 0) i0 = 0;
      else i0 = (num + 2 * dy - 1) / (2 * dy);
      if (i0 > nx) i0 = nx;
      /* an edge heading +y has the +x side on its right in the raw plane: samples to the +x side
       * of an upward edge get -1 ... with counter-clockwise = positive area for
       * sum (x_i y_{i+1} - x_{i+1} y_i): a ccw square's right side heads +y and the inside is on
       * its -x side, so crossing it towards +x LEAVES the region: -1 for upward edges */
      diff[i0] -= sign;
    }
    int32_t w = 0;
    /* winding left of every crossing is 0 (outside the bounding box) */
    for (int i = 0; i < nx; ++i) {
      w += diff[i];
      out[(int64_t)k * nx + i] = (int16_t)w;
    }
  }
  free(diff);
}
